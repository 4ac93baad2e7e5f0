"""
Minimal text log with the interface JetModel/Pipeline use in the reference
(logger/logger.py:63-150: ``Log(fname, verbose)``, ``add_entry(mtype, entry,
timestamp)``, ``entries``).  Out of the hot path; kept so `log.add_entry(...)` calls
made by reference-side callers keep working against the new JetModel.
"""
import os
import time

VALID_MTYPES = ("INFO", "ERROR", "WARNING")


class Entry:
    def __init__(self, mtype, entry, timestamp=True):
        if not isinstance(mtype, str):
            raise TypeError("mtype must be a str")
        if not isinstance(entry, str):
            raise TypeError("entry must be a str")
        if mtype.upper() not in VALID_MTYPES:
            raise TypeError("mtype must be one of '" + "', '".join(VALID_MTYPES) + "'")
        self.mtype = mtype.upper()
        self.message = entry
        self.rtime = time.localtime() if timestamp else None

    def __setstate__(self, state):
        # entries pickled by the reference (logger/logger.py:180-209) carry _mtype / _message /
        # _mtime / timestamp
        if '_message' in state:
            self.mtype = str(state.get('_mtype', 'INFO')).upper()
            self.message = state['_message']
            self.rtime = state.get('_mtime') if state.get('timestamp', True) else None
        else:
            self.__dict__.update(state)

    def __str__(self):
        stamp = time.strftime("%d%b%Y-%H:%M:%S", self.rtime).upper() if self.rtime else ""
        pre = "::".join([s for s in (stamp, self.mtype.ljust(7)) if s])
        return ": ".join([pre, self.message])


class Log:
    def __init__(self, fname, verbose=True):
        self._filename = fname
        self._verbose = verbose
        self._entries = []
        dcy = os.path.dirname(os.path.abspath(fname))
        if not os.path.isdir(dcy):
            raise FileNotFoundError(f"{dcy} is not a directory")

    def __setstate__(self, state):
        self.__dict__.update(state)
        if isinstance(self._entries, dict):     # the reference numbers its entries in a dict
            self._entries = [self._entries[k] for k in sorted(self._entries)]

    def __str__(self):
        return "\n".join(str(e) for e in self._entries)

    @property
    def filename(self):
        return self._filename

    @property
    def verbose(self):
        return self._verbose

    @verbose.setter
    def verbose(self, new_verbosity):
        self._verbose = bool(new_verbosity)

    @property
    def entries(self):
        return self._entries

    def add_entry(self, mtype, entry, timestamp=True):
        e = Entry(mtype, entry, timestamp)
        self._entries.append(e)
        with open(self._filename, "at") as f:
            f.write(str(e) + "\n")
        if self._verbose:
            print(str(e))
