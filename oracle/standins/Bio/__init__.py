"""Minimal stand-in for Biopython (TEST INFRASTRUCTURE ONLY)."""
__version__ = "0.0-standin"
