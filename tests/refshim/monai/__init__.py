"""Empty stand-in for `import monai` (src/py/utils.py:15); see tests/refshim/vtk."""
