"""ORACLE — CPU restatement of the reference path; test infrastructure only (see drn_oracle.py)."""
