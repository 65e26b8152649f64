"""CPU oracle of the MeshODE hot path -- TEST INFRASTRUCTURE ONLY (see meshode_oracle.cc)."""
