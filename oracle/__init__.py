"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU checkers (see oracle/ntt_oracle.h)."""
