"""Stand-in for the un-vendored third-party dependency ``kornia==0.6.3``.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The reference imports exactly one
symbol from kornia (``/root/reference/scripts/homography.py:2``; pin in
``/root/reference/requirements.txt:1``).  kornia is not installed in this image and
there is no network, so this package restates the published algorithm of
``kornia.geometry.transform.warp_perspective`` (0.6.3) so that the *unmodified*
reference can be imported and used to generate golden vectors.
"""
__version__ = "0.6.3+standin"
from . import geometry  # noqa: F401
