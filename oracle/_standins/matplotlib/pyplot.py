"""Empty stand-in for matplotlib.pyplot (never called on the hot path)."""
