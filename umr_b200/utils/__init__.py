"""Drop-in counterparts of the reference's `utils/` helpers that sit on the per-step path."""
