"""ORACLE package -- test infrastructure only (see maskbit_oracle.py / select_oracle.c headers)."""
