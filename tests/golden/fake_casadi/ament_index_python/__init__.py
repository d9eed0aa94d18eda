"""Stub of ament_index_python for tests/golden/make_golden.py (build container only)."""
