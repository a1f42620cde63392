"""Host-side mirror of the reference's ``src`` package (code objects and chains)."""
