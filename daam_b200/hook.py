"""Hook plumbing: context-managed monkey patching and the walker that finds a UNet's cross-attention layers.

Behavioural mirror of the reference's L0 (``/root/reference/daam/hook.py``): same class and method names, same
exceptions (``RuntimeError('Already hooked module')`` / ``('Module is not hooked')``, hook.py:36-37, 46-47), and --
what the hot path depends on -- the same layer enumeration order, which defines ``layer_idx`` (hook.py:95-127):
``up_blocks`` first, then ``down_blocks``, then optionally ``mid_block``.
"""
from __future__ import annotations

import functools
from typing import Generic, List, Optional, Set, TypeVar

import torch.nn as nn

__all__ = ['ObjectHooker', 'ModuleLocator', 'AggregateHooker', 'UNetCrossAttentionLocator']

ModuleType = TypeVar('ModuleType')
ModuleListType = TypeVar('ModuleListType', bound=List)


class ModuleLocator(Generic[ModuleType]):
    def locate(self, model: nn.Module) -> List[ModuleType]:
        raise NotImplementedError


class ObjectHooker(Generic[ModuleType]):
    """Owns one object; ``hook()`` applies patches, ``unhook()`` restores every attribute ``monkey_patch`` replaced."""

    def __init__(self, module: ModuleType):
        self.module: ModuleType = module
        self.hooked = False
        self._originals = {}

    def __enter__(self):
        self.hook()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.unhook()

    def hook(self):
        if self.hooked:
            raise RuntimeError('Already hooked module')
        self._originals = {}
        self.hooked = True
        self._hook_impl()
        return self

    def unhook(self):
        if not self.hooked:
            raise RuntimeError('Module is not hooked')
        for name, fn in self._originals.items():
            setattr(self.module, name, fn)
        self.hooked = False
        self._unhook_impl()
        return self

    def monkey_patch(self, fn_name: str, fn, strict: bool = True):
        """Replace ``module.fn_name`` by ``fn(module, ...)``; a missing attribute is ignored unless ``strict``."""
        try:
            self._originals[fn_name] = getattr(self.module, fn_name)
        except AttributeError:
            if strict:
                raise
            return
        setattr(self.module, fn_name, functools.partial(fn, self.module))

    def monkey_super(self, fn_name: str, *args, **kwargs):
        return self._originals[fn_name](*args, **kwargs)

    def _hook_impl(self):
        raise NotImplementedError

    def _unhook_impl(self):
        pass


class AggregateHooker(ObjectHooker[ModuleListType]):
    """A hooker over a list of hookers."""

    def _hook_impl(self):
        for child in self.module:
            child.hook()

    def _unhook_impl(self):
        for child in self.module:
            child.unhook()

    def register_hook(self, hook: ObjectHooker):
        self.module.append(hook)


class UNetCrossAttentionLocator(ModuleLocator):
    """Enumerates ``attn2`` modules in the reference's order; the position in the returned list is ``layer_idx``."""

    def __init__(self, restrict: Optional[Set[int]] = None, locate_middle_block: bool = False):
        self.restrict = restrict
        self.layer_names: List[str] = []
        self.locate_middle_block = locate_middle_block

    def locate(self, model) -> list:
        self.layer_names.clear()
        located = []
        tagged = [(blk, 'up') for blk in model.up_blocks] + [(blk, 'down') for blk in model.down_blocks]
        if self.locate_middle_block:
            tagged.append((model.mid_block, 'mid'))
        for block, tag in tagged:
            if 'CrossAttn' not in type(block).__name__:
                continue
            layers = [tb.attn2 for transformer in block.attentions for tb in transformer.transformer_blocks]
            for i, layer in enumerate(layers):      # the index restarts in every block: names are not unique
                if self.restrict is None or i in self.restrict:
                    located.append(layer)
                    self.layer_names.append(f'{tag}-attn-{i}')
        return located
