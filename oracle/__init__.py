"""CPU oracle (test infrastructure only) -- see hqq_oracle.py."""
