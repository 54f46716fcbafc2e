"""Small helpers kept from the reference's API surface (``/root/reference/daam/utils.py``): seeding, device/autocast
shims and the word -> token-row lookup (``compute_token_merge_indices``, utils.py:73-91). spaCy and plotting helpers
are out of scope."""
from __future__ import annotations

import os
import random
import sys
from pathlib import Path
from typing import List, Optional, Tuple, TypeVar

import numpy as np
import torch

__all__ = ['set_seed', 'compute_token_merge_indices', 'cache_dir', 'auto_device', 'auto_autocast']

T = TypeVar('T')


def auto_device(obj: T = torch.device('cpu')) -> T:
    """``torch.device`` -> the best device; anything else -> moved to CUDA when there is one (utils.py:22-29)."""
    has_cuda = torch.cuda.is_available()
    if isinstance(obj, torch.device):
        return torch.device('cuda' if has_cuda else 'cpu')
    return obj.to('cuda') if has_cuda else obj


def auto_autocast(*args, **kwargs):
    """``torch.autocast('cuda', ...)`` that switches itself off without a GPU (utils.py:32-36)."""
    if not torch.cuda.is_available():
        kwargs['enabled'] = False
    return torch.autocast('cuda', *args, **kwargs)


def set_seed(seed: int) -> torch.Generator:
    """Seeds python, numpy and torch (all devices) and returns a seeded generator on the auto device (utils.py:46-55)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    gen = torch.Generator(device=auto_device())
    gen.manual_seed(seed)
    return gen


def cache_dir() -> Path:
    """Per-user cache folder, honouring XDG_CACHE_HOME on Linux (utils.py:58-70)."""
    if sys.platform == 'darwin':
        return Path(os.path.expanduser('~'), 'Library/Caches/daam')
    if os.name == 'posix':
        return Path(os.environ.get('XDG_CACHE_HOME', os.path.expanduser('~/.cache')), 'daam')
    return Path(os.environ.get('LOCALAPPDATA') or os.path.expanduser('~\\AppData\\Local'), 'daam')


def _pieces(tokenizer, text: str) -> List[str]:
    return [tok.replace('</w>', '') for tok in tokenizer.tokenize(text)]


def compute_token_merge_indices(tokenizer, prompt: str, word: str, word_idx: Optional[int] = None,
                                offset_idx: int = 0) -> Tuple[List[int], Optional[int]]:
    """Rows of the global heat map that belong to ``word``: every occurrence of the word's token pieces in the
    lower-cased prompt, shifted by one for the SOS row. With ``word_idx`` the lookup is skipped and ``[word_idx + 1]``
    returned. Raises ``ValueError('Search word ... not found in prompt!')`` like utils.py:86-87."""
    if word_idx is not None:
        return [word_idx + 1], word_idx
    haystack = _pieces(tokenizer, prompt.lower())
    word = word.lower()
    needle = _pieces(tokenizer, word)
    n = len(needle)
    rows: List[int] = []
    for start in range(len(haystack)):
        if haystack[start:start + n] == needle:
            rows.extend(start + offset_idx + 1 + j for j in range(n))
    if not rows:
        raise ValueError(f'Search word {word} not found in prompt!')
    return rows, word_idx
