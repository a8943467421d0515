"""Empty stand-in: utils/utilities.py imports h5py at module scope; nothing on the tested path uses it."""
