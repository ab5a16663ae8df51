"""CPU oracle for the PIML hot path -- TEST INFRASTRUCTURE ONLY (see piml_oracle.c). Never imported by piml_b200."""
