"""TEST INFRASTRUCTURE ONLY -- no-op `matplotlib` (plots are out of scope; never pass --plot)."""
