"""Host-side mirror of the reference's `utils` package (hot-path members only)."""
