"""CPU oracles for the CUDA hot path -- test infrastructure only (never imported by the product package)."""
