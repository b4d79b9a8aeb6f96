"""The reference's labels.json:1-31 (29 symbols, blank '_' at index 0)."""
LABELS = ["_", "'"] + [chr(c) for c in range(ord('A'), ord('Z') + 1)] + [" "]
