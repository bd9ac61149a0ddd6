"""Drop-in mirror of the reference's ``mapping`` package (GridMap / Mapping) on the B200 engine."""
