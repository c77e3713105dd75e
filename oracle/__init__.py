"""CPU oracle — test infrastructure only (see oracle/gab1_oracle.c).  Never imported by the product."""
