"""CPU oracle for the frame-producing hot path (test infrastructure; see analyser_oracle.py)."""
