"""ntracer_b200.wrapper -- NTracer(dimension): the helper of the reference's lib/ntracer/wrapper.py that fills in
the `dimension` argument.  Instances are cached per dimension; `force_generic` is accepted for compatibility (the
backend always picks the fixed-dimension kernels for 3..8 and the run-time-dimension kernels otherwise)."""
import weakref

from . import tracern

CUBE = 1
SPHERE = 2


class NTracer:
    _cache = weakref.WeakValueDictionary()

    def __new__(cls, dimension, force_generic=False):
        if not force_generic:
            obj = NTracer._cache.get(dimension)
            if obj is not None:
                return obj
        tracern._check_dimension(dimension)
        obj = object.__new__(cls)
        dim = dimension
        mod = tracern
        obj.dimension = dim
        obj.base = mod

        # Like the reference's wrappers (lib/ntracer/wrapper.py:10-66) the curried classes hand back instances of the
        # BASE classes -- `nt.Vector(1,2,3)` is a tracern.Vector -- so objects made through different NTracer instances
        # mix freely and pickle by their module-level type.
        def curried(base, make, **statics):
            ns = {'__slots__': (), '__new__': lambda cls, *a, **k: make(*a, **k)}
            ns.update({k: staticmethod(v) for k, v in statics.items()})
            return type(base.__name__, (base,), ns)

        def spread(base):
            # Vector / Matrix accept their values as separate arguments as well as one sequence
            return lambda *values: base(dim, values) if len(values) > 1 else base(dim, *values)

        Vector = curried(mod.Vector, spread(mod.Vector), axis=lambda axis, length=1: mod.Vector.axis(dim, axis, length))
        Matrix = curried(mod.Matrix, spread(mod.Matrix),
                         scale=lambda factor: mod.Matrix.scale(factor) if isinstance(factor, mod.Vector) else mod.Matrix.scale(dim, factor),
                         identity=lambda: mod.Matrix.identity(dim))
        Camera = curried(mod.Camera, lambda: mod.Camera(dim))
        BoxScene = curried(mod.BoxScene, lambda: mod.BoxScene(dim))
        AABB = curried(mod.AABB, lambda *a, **k: mod.AABB(dim, *a, **k))
        obj.Vector, obj.Matrix, obj.Camera, obj.BoxScene, obj.AABB = Vector, Matrix, Camera, BoxScene, AABB
        for n in ['CompositeScene', 'KDNode', 'KDLeaf', 'KDBranch', 'Primitive', 'PrimitiveBatch', 'PrimitivePrototype',
                  'Solid', 'SolidPrototype', 'Triangle', 'TriangleBatch', 'TrianglePrototype', 'TriangleBatchPrototype',
                  'PointLight', 'GlobalLight', 'dot', 'cross', 'build_kdtree', 'build_composite_scene',
                  'screen_coord_to_ray', 'BATCH_SIZE']:
            setattr(obj, n, getattr(mod, n))
        if not force_generic:
            NTracer._cache[dimension] = obj
        return obj
