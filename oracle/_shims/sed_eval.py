"""Empty stand-in: utils/utilities.py imports sed_eval at module scope; nothing on the tested path uses it."""
