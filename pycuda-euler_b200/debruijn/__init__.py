"""debruijn package of the reference layout (src/debruijn/)."""
