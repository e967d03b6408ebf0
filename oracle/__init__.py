"""CPU oracle for stage 1 -- TEST INFRASTRUCTURE ONLY (see oracle/stage1_oracle.c)."""
