"""TEST INFRASTRUCTURE ONLY: CPU oracle for the D2Q9 path (see oracle/d2q9_oracle.c)."""
