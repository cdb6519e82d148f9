"""Test infrastructure only: CPU oracles (never imported by the product path)."""
