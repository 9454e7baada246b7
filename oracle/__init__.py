"""oracle/ — CPU restatement of the reference's hot path.  Test infrastructure only: nothing
under fedmlp_b200/ may import this package (see oracle/fedmlp_oracle.py header)."""
