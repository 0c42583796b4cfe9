"""CPU oracle for the kameris hot paths -- TEST INFRASTRUCTURE ONLY (see kameris_oracle.c)."""
