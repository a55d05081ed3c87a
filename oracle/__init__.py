"""CPU oracle (test infrastructure only - see newtonnet_oracle.py)."""
