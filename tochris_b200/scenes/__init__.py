"""Procedural scene generators (palette, colormap, BSP29, MDL6, camera sets) in the
reference's on-disk formats -- SURVEY.md 8(d), appendix B."""
