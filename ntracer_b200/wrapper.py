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

        class Vector(mod.Vector):
            __slots__ = ()

            def __init__(self, *values):
                if len(values) > 1:
                    mod.Vector.__init__(self, dim, values)
                else:
                    mod.Vector.__init__(self, dim, *values)

            @staticmethod
            def axis(axis, length=1):
                return mod.Vector.axis(dim, axis, length)

        class Matrix(mod.Matrix):
            __slots__ = ()

            def __init__(self, *values):
                if len(values) > 1:
                    mod.Matrix.__init__(self, dim, values)
                else:
                    mod.Matrix.__init__(self, dim, *values)

            @staticmethod
            def scale(factor):
                if isinstance(factor, mod.Vector):
                    return mod.Matrix.scale(factor)
                return mod.Matrix.scale(dim, factor)

            @staticmethod
            def identity():
                return mod.Matrix.identity(dim)

        class Camera(mod.Camera):
            def __init__(self):
                mod.Camera.__init__(self, dim)

        class BoxScene(mod.BoxScene):
            def __init__(self):
                mod.BoxScene.__init__(self, dim)

        class AABB(mod.AABB):
            def __init__(self, *args, **kwds):
                mod.AABB.__init__(self, dim, *args, **kwds)

        obj.Vector, obj.Matrix, obj.Camera, obj.BoxScene, obj.AABB = Vector, Matrix, Camera, BoxScene, AABB
        for n in ['CompositeScene', 'KDNode', 'KDLeaf', 'KDBranch', 'Primitive', 'PrimitiveBatch', 'PrimitivePrototype',
                  'Solid', 'SolidPrototype', 'Triangle', 'TriangleBatch', 'TrianglePrototype', 'TriangleBatchPrototype',
                  'PointLight', 'GlobalLight', 'dot', 'cross', 'build_kdtree', 'build_composite_scene',
                  'screen_coord_to_ray', 'BATCH_SIZE']:
            setattr(obj, n, getattr(mod, n))
        if not force_generic:
            NTracer._cache[dimension] = obj
        return obj
