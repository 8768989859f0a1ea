"""Index helper used by ``QuantLinear.set_kernel`` for o_proj (reference: qeft/reorder.py:6-12)."""
import torch


def sparse_to_dense_ids(sparse_ids: torch.Tensor, length: int) -> torch.Tensor:
    """Permutation that moves the columns listed in ``sparse_ids`` to the end, keeping the rest in order."""
    if not len(sparse_ids) < length:
        raise AssertionError("outlier index list must be shorter than the feature dimension")
    ids = sparse_ids.to(torch.long)
    keep = torch.ones(length, dtype=torch.bool, device=ids.device)
    keep[ids] = False
    rest = torch.nonzero(keep, as_tuple=False).flatten()
    return torch.cat([rest, ids])
