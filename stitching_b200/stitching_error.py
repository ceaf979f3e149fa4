"""Error conventions of the boundary (mirrors stitching/stitching_error.py:1-6)."""


class StitchingError(Exception):
    pass


class StitchingWarning(UserWarning):
    pass
