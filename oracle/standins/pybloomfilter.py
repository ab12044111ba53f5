"""Stand-in for `pybloomfilter` (TEST INFRASTRUCTURE ONLY): an exact, file-backed set.

Deviation from the real library: no hash false positives."""
import os
import pickle


class BloomFilter:
    def __init__(self, capacity=0, error_rate=0.05, filename=None):
        self._items = set()
        self._filename = filename
        self._readonly = False

    @classmethod
    def open(cls, filename, mode="rw"):
        bf = cls(filename=filename)
        with open(filename, "rb") as fh:
            bf._items = pickle.load(fh)
        bf._readonly = "w" not in mode
        return bf

    def add(self, key):
        self._items.add(key)

    def __contains__(self, key):
        return key in self._items

    def sync(self):
        if self._filename and not self._readonly:
            tmp = self._filename + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as fh:
                pickle.dump(self._items, fh)
            os.replace(tmp, self._filename)

    def close(self):
        self.sync()
