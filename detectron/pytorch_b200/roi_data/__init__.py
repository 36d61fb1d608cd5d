"""Device-side counterparts of the reference's lib/roi_data."""
