"""CPU oracle package -- test infrastructure only (see gemlite_oracle.py header)."""
