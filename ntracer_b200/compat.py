"""Pickles that cross between this package and the reference.

Both sides reduce their value types to the same functions with the same payload (ntracer_b200/render.py: the
reference's encodings, src/render.cpp:1391-1657); what differs is the module the functions are looked up in
(`ntracer.render` there, `ntracer_b200.render` here).  `loads_reference` reads a pickle the reference wrote,
`dumps_for_reference` writes one the reference can read."""
import io
import pickle

_TO_HERE = {'ntracer.render': 'ntracer_b200.render', 'ntracer.tracern': 'ntracer_b200.tracern', 'ntracer.wrapper': 'ntracer_b200.wrapper'}
_TO_REFERENCE = {v: k for k, v in _TO_HERE.items()}


class ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module in _TO_HERE or module.startswith('ntracer.tracer'):      # ntracer.tracer3 ... tracer8: the per-dimension modules
            module = _TO_HERE.get(module, 'ntracer_b200.tracern')
        return super().find_class(module, name)


def load_reference(file):
    return ReferenceUnpickler(file).load()


def loads_reference(data):
    return load_reference(io.BytesIO(data))


class ReferencePickler(pickle._Pickler):
    """The pure-Python pickler with the module of this package's unpickle functions written as the reference's."""
    def save_global(self, obj, name=None):
        module = getattr(obj, '__module__', None)
        if module in _TO_REFERENCE and name is None:
            name = getattr(obj, '__qualname__', obj.__name__)
            target = _TO_REFERENCE[module]
            if self.proto >= 4:
                self.save(target)
                self.save(name)
                self.write(pickle.STACK_GLOBAL)
            else:
                self.write(pickle.GLOBAL + target.encode('ascii') + b'\n' + name.encode('ascii') + b'\n')
            self.memoize(obj)
            return
        super().save_global(obj, name)

    dispatch = dict(pickle._Pickler.dispatch)
    import types as _types
    dispatch[_types.FunctionType] = save_global


def dump_for_reference(obj, file, protocol=2):
    ReferencePickler(file, protocol).dump(obj)


def dumps_for_reference(obj, protocol=2):
    f = io.BytesIO()
    dump_for_reference(obj, f, protocol)
    return f.getvalue()
