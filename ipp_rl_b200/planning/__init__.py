"""Drop-in mirror of the hot-path part of the reference's ``planning`` package."""
