"""CPU oracle -- TEST INFRASTRUCTURE.  See oracle/clane_oracle.c and oracle/oracle.py."""
