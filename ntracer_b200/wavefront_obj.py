"""Wavefront .obj reader next to the path (the reference ships lib/ntracer/wavefront_obj.py with the
same interface): vertices from `v x y z [w]` lines, polygons from `f` lines (`i`, `i/t`, `i/t/n`, `i//n`; 1-based, or
negative = relative to the vertices read so far) split into triangle fans, everything else ignored.
load_obj(file[, nt]) -> list of TrianglePrototype with Material((1,1,1)), ready for build_composite_scene."""
from . import render
from . import wrapper


class FileFormatError(Exception):
    def __init__(self):
        super().__init__('not a valid wavefront file')


def _vertex_index(token, count):
    """index field of a face corner -> 0-based position in the vertex list read so far"""
    i = int(token.split('/', 1)[0], 10)
    if i > 0:
        i -= 1
    elif i < 0:
        i += count
    else:
        raise IndexError(token)
    if not 0 <= i < count:
        raise IndexError(token)
    return i


def load_obj(file, nt=None):
    if nt is None:
        nt = wrapper.NTracer(3)
    elif nt.dimension != 3:
        raise ValueError('Wavefront .obj files only support 3-dimensional geometry')
    material = render.Material((1, 1, 1))
    vertices, prototypes = [], []
    with open(file, 'r') as src:
        for line in src:
            fields = line.split()
            if not fields:
                continue
            try:
                if fields[0] == 'v':
                    xyz = [float(x) for x in fields[1:4]]
                    if len(xyz) != 3:
                        raise ValueError(line)
                    vertices.append(nt.Vector(xyz))
                elif fields[0] == 'f':
                    corners = [vertices[_vertex_index(t, len(vertices))] for t in fields[1:]]
                    for k in range(1, len(corners) - 1):
                        prototypes.append(nt.TrianglePrototype([corners[0], corners[k], corners[k + 1]], material))
            except (ValueError, IndexError):
                raise FileFormatError() from None
    return prototypes
