"""Host utilities of the fit drivers (mirror of the pieces of mufasa/utils the hot path touches)."""
