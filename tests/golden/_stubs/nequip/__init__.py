"""Stand-in for nequip (absent here); see ../README.md."""
