"""TEST INFRASTRUCTURE: CPU oracles for the BM25 / cosine hot path (checker + CPU baseline only)."""
