"""CPU oracle for the face-generator hot path. TEST INFRASTRUCTURE ONLY (see fg_oracle.cpp)."""
