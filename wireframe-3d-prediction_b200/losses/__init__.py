"""Drop-in `losses` package (see losses/WireframeLoss.py)."""
