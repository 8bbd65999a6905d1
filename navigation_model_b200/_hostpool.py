"""A few host threads for the bulk numpy work around a launch (range checks, biomechanical-cost tables, copies into
and out of pinned staging memory): numpy releases the GIL inside its loops.  Created on first use, shared by the
package; the device work itself never goes through it."""
import os
import threading
from concurrent.futures import ThreadPoolExecutor

_POOL = None
_LOCK = threading.Lock()


def workers():
    return max(1, min(16, os.cpu_count() or 1))


def pool():
    global _POOL
    if _POOL is None:
        with _LOCK:
            if _POOL is None:
                _POOL = ThreadPoolExecutor(max_workers=workers(), thread_name_prefix="nm-host")
    return _POOL
