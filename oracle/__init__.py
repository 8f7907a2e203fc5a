"""CPU oracle package - test infrastructure only (see ref_torch.py header)."""
