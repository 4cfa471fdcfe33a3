"""Host-side persistence around the ITE step: the reference pickles the unit cell after EVERY edge update
(src/algo/imaginary_time_evolution/_tn_update.py:203 -> src/unit_cell/definition.py:119-124 -> src/utils/saveload.py:143-172),
synchronously, on the thread that drives the step.  Once the step itself takes tens of milliseconds that write is a visible
share of it, so here the same files are written by ONE background thread fed from a queue: the step hands over the (fresh,
never mutated again) host arrays and goes on; ``flush()`` waits for the disk.

File format: a pickle of a plain dict  {"format": "kbp-unit-cell-v1", "A", "B", "C", "file_name"}  -- no class of this
package or of the reference is needed to read it; ``UnitCell.load`` accepts it, and also a pickle of any object with
``A / B / C`` attributes (the reference's own ``UnitCell`` pickles, when its modules are importable).
"""
from __future__ import annotations

import os
import pickle
import queue
import threading
import time

import numpy as np

DEFAULT_FOLDER = os.path.join(os.getcwd(), "data", "unit_cells")


def time_stamp() -> str:
    """(src/utils/strings.py time_stamp) yyyy.mm.dd_hh.mm.ss + a random suffix against collisions"""
    t = time.localtime()
    return time.strftime("%Y.%m.%d_%H.%M.%S", t) + "_" + "".join(np.random.choice(list("ABCDEFGHIJKLMNOPQRSTUVWXYZ"), 5))


def unit_cell_payload(unit_cell, file_name=None) -> dict:
    return {"format": "kbp-unit-cell-v1", "A": np.array(unit_cell.A), "B": np.array(unit_cell.B), "C": np.array(unit_cell.C),
            "file_name": file_name}


def write_payload(payload: dict, fullpath: str) -> str:
    os.makedirs(os.path.dirname(fullpath) or ".", exist_ok=True)
    tmp = fullpath + ".tmp"
    with open(tmp, "wb") as f:
        pickle.dump(payload, f, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(tmp, fullpath)                 # a reader never sees a half-written file
    return fullpath


def read_unit_cell_arrays(fullpath: str):
    with open(fullpath, "rb") as f:
        obj = pickle.load(f)
    if isinstance(obj, dict):
        return np.asarray(obj["A"]), np.asarray(obj["B"]), np.asarray(obj["C"])
    return np.asarray(obj.A), np.asarray(obj.B), np.asarray(obj.C)


class AsyncSaver:
    """one writer thread; ``submit`` never blocks on the disk."""

    def __init__(self):
        self.q: queue.Queue = queue.Queue()
        self.errors: list = []
        self.written = 0
        self._t = threading.Thread(target=self._run, name="kbp-saver", daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self.q.get()
            try:
                if item is None:
                    return
                payload, path = item
                write_payload(payload, path)
                self.written += 1
            except Exception as e:              # surfaced by flush(): a failed save must not be silent
                self.errors.append(e)
            finally:
                self.q.task_done()

    def submit(self, payload: dict, fullpath: str):
        self.q.put((payload, fullpath))

    def flush(self):
        self.q.join()
        if self.errors:
            e, self.errors = self.errors[0], []
            raise e

    def close(self):
        self.flush()
        self.q.put(None)
        self._t.join(timeout=5)


_saver = None


def saver() -> AsyncSaver:
    global _saver
    if _saver is None:
        _saver = AsyncSaver()
    return _saver
