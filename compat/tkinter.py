"""scene/deformation.py:5 does `from tkinter import W` (stray IDE import)."""
W = "w"
