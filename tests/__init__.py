"""Parity tests of spfsplatv2_b200 (CPU: oracle / golden / ABI; GPU: CUDA path vs oracle)."""
