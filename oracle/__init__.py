"""Test-only CPU oracle for the SMPLify / SMPL hot path (see oracle/port.py)."""
