"""oracle/ — TEST INFRASTRUCTURE ONLY (checker for the CUDA path; never imported by sciport_rs_b200)."""
