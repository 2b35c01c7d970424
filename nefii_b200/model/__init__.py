"""Host-side mirror of the reference's ``code/model`` package (same names, same signatures)."""
