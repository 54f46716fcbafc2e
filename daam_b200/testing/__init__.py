"""Test and benchmark fixtures -- NOT part of the product path.

``synthetic.py`` restates the diffusers 0.21.2 surface the tracer hooks into (``Attention`` with ``set_processor``,
a ``UNet2DConditionModel``-shaped module tree, a ``StableDiffusionPipeline``-shaped driver) with random-init weights,
because neither diffusers nor any checkpoint is available offline. ``tests/``, ``bench.py``, ``__graft_entry__.smoke()``
and the oracle's reference loader build their pipelines from it; nothing under ``daam_b200/*.py`` imports it.
"""
