"""Mask helpers with the reference's names and semantics (reference utils.py:11-37, 47-70).
The CUDA attention kernels never materialise these masks (they read token ids directly); the
functions exist for callers of the reference API and for tests."""
import torch


def generate_square_subsequent_mask(sz: int, device="cpu") -> torch.Tensor:
    """Float (sz, sz): 0 where key <= query, -inf strictly above the diagonal."""
    future = torch.ones(sz, sz, dtype=torch.bool, device=device).triu(diagonal=1)
    return torch.zeros(sz, sz, device=device).masked_fill(future, float("-inf"))


def create_padding_mask(seq: torch.Tensor, pad_idx: int = 0) -> torch.Tensor:
    """Bool (B, T), True at padding positions.  Stays on seq's device (the reference moves it to
    the global config.DEVICE, reference utils.py:70, which breaks CPU tensors on a CUDA box)."""
    return seq == pad_idx
