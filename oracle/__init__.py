"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see femx_oracle.c header)."""
