"""Model-level callers of the hot path: walker, quantised-checkpoint save/load."""
