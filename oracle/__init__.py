"""TEST INFRASTRUCTURE ONLY: CPU oracle for the fusion path (see oracle/emap_oracle.c)."""
