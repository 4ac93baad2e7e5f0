"""
Loading of save files written by the REFERENCE (`JetModel.save`, classes.py:1704-1713;
`Pipeline.save`, :2215-2258): their pickles name classes of the package `RaJePy`
(`RaJePy.logger.logger.Log` / `Entry`, `RaJePy.classes.ContinuumRun` / `RRLRun`).  The
unpickler below maps those names onto this package's equivalents, whose attribute names are
the reference's, so `-r/--resume` keeps working on directories the reference produced.
"""
import io
import pickle


class _ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        root = module.split('.')[0]
        if root == 'RaJePy':
            from . import logger, pipeline
            table = {'Log': logger.Log, 'Entry': logger.Entry,
                     'ContinuumRun': pipeline.ContinuumRun, 'RRLRun': pipeline.RRLRun}
            if name in table:
                return table[name]
            raise pickle.UnpicklingError(f"save file refers to {module}.{name}, which has no "
                                         f"equivalent in rajepy_b200")
        return super().find_class(module, name)


def load_pickle(path):
    with open(path, 'rb') as f:
        return _ReferenceUnpickler(io.BytesIO(f.read())).load()
