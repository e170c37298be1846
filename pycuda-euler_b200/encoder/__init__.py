"""encoder package of the reference layout (src/encoder/)."""
