"""Import the reference's own leaf modules from /root/reference (build container only).

The reference is pure Python; its hot-path primitives import once ``pyvips`` is
stubbed.  ``traiNNer.models`` cannot be imported (spandrel /
ema_pytorch missing, hard-coded .cuda()), so only the primitives are loaded.
Nothing here runs on the GPU box: /root/reference does not exist there.
"""

from __future__ import annotations

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("OTF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "traiNNer"))


def load_reference(seed: int = 0) -> SimpleNamespace:
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.modules.setdefault("pyvips", MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from traiNNer.utils.rng import RNG  # noqa: PLC0415

    if RNG._rng is None:
        RNG.init_rng(seed)
    from traiNNer.data import degradations as deg  # noqa: PLC0415
    from traiNNer.data import transforms as tfm  # noqa: PLC0415
    from traiNNer.utils import diffjpeg as dj  # noqa: PLC0415
    from traiNNer.utils import img_process_util as ipu  # noqa: PLC0415

    return SimpleNamespace(deg=deg, tfm=tfm, dj=dj, ipu=ipu, RNG=RNG)
