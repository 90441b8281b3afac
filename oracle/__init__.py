"""CPU oracle package (test infrastructure only; see quilt_oracle.cpp)."""
