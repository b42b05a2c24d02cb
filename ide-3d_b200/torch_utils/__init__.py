"""Drop-in counterpart of the reference's `torch_utils` package (only what the hot path touches)."""
