"""gpuhash package of the reference layout (src/gpuhash/)."""
