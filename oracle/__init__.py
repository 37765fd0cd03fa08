"""CPU oracle of the S^3 hot path -- TEST INFRASTRUCTURE ONLY (see oracle/s3_oracle.py)."""
