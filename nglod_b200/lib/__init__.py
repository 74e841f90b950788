"""Host-side mirror of the reference's `sdf-net/lib` package for the hot path."""
