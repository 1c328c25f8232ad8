"""TEST INFRASTRUCTURE ONLY (see oracle/rq_oracle.py header).  Never imported by the product package."""
