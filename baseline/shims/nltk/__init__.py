"""Stand-in for `nltk` (not installed; only the reference's datasets.py imports it, for caption tokenising)."""
