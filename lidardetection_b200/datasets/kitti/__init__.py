"""Only the evaluation-side callers of the rotated-box hot path live here; datasets themselves are out of scope."""
