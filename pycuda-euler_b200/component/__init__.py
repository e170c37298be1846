"""component package of the reference layout (src/component/)."""
