"""Test-only CPU oracle (see gp_oracle.py header). Not part of the product."""
