"""CPU oracle for the ros2_mpc NMPC solve — TEST INFRASTRUCTURE ONLY (see mpc_oracle.h)."""
