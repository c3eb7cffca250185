"""numbskull_b200 (package init filled in later)."""
