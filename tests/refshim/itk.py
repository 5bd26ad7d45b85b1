"""Empty stand-in: the reference's utils.py imports itk at module level (src/py/utils.py:7);
nothing the golden generator executes touches it."""
