"""eulertour package of the reference layout (src/eulertour/)."""
