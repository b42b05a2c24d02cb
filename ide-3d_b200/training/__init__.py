"""Counterpart of the reference's `training` package: volumetric_rendering (free functions), networks (StyleGAN2
blocks), triplane (the generator the checkpoints would carry)."""
