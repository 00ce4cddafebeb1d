"""Empty import stub: lets the read-only reference import in a container without this package (golden-vector generation only)."""
