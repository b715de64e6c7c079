"""Seed control for the counter-based per-chain RNG that replaces the reference's global MT19937
(sampling.pyx:425-447).  ``set_seed`` has the reference's meaning -- it fixes every random choice that follows --
but there is no global stream to advance: the kernels derive an independent Philox stream from
(seed, global chain id, sweep index, event id), see ``csrc/bq_rng.cuh``.
"""
import numpy

_seed = 13  # sampling.pyx:438
_epoch = 0  # bumped per engine so that two engines built under one seed do not replay each other


def set_seed(seed):
    global _seed, _epoch
    _seed = int(seed)
    _epoch = 0
    numpy.random.seed(int(seed) % (2 ** 32))


def get_seed():
    return _seed


def next_engine_seed():
    """Seed handed to the next GibbsEngine: (user seed, engine ordinal) mixed to 64 bits."""
    global _epoch
    s = (_seed * 0x9E3779B97F4A7C15 + _epoch * 0xBF58476D1CE4E5B9 + 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    _epoch += 1
    return s
