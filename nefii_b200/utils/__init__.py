"""Host-side helpers mirroring the reference's ``code/utils`` functions that the hot path calls."""
