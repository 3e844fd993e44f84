"""Batch sharding for counterfactual sampling: independent (image, intervention) units split across ranks with no
data-path collective until the final gather (ref scripts/image_causaldae_test.py:438-440)."""


def shard_range(n_units, rank, world):
    """contiguous, ordered, covering split of range(n_units); sizes differ by at most one"""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
