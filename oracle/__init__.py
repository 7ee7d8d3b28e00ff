"""CPU oracle for the tracking hot path -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py)."""
