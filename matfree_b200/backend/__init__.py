"""Thin host-side helpers (device arrays, PRNG keys).  Not a multi-backend layer."""
