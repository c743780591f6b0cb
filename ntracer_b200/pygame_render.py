"""ntracer_b200.pygame_render -- the reference's Pygame front end (lib/ntracer/pygame_render.py:8-127) on this backend:
`channels_from_surface(surface)` describes a Surface's pixel layout as render.Channel objects, `PygameRenderer` is a
CallbackRenderer that draws onto a Surface and posts a Pygame event when the frame is complete (the loop of
scripts/polytope.py:505-557).  Needs pygame, like the reference's module; nothing else in the package imports it."""
import weakref

import pygame

from . import render


def channels_from_surface(surface):
    """The list of render.Channel objects that matches the pixel format of `surface` (lib/ntracer/pygame_render.py:8-46).

    ImageFormat counts a pixel's bits from the most significant end of its `bytesize` bytes, Pygame gives every colour
    component as (loss, shift, mask) counted from the least significant end; unused bit ranges become channels whose
    factors are all zero.  Indexed (8-bit) modes are not supported."""
    nbytes = surface.get_bytesize()
    if nbytes == 1:
        raise TypeError('indexed color modes are not supported')
    top = (nbytes - 1) * 8
    fields = []
    for name, loss, shift in zip('RGBA', surface.get_losses(), surface.get_shifts()):
        bits = 8 - loss
        if bits:
            fields.append((top + loss - shift, bits, name))     # first bit of the component, counted from the top
    fields.sort()
    channels, pos = [], 0
    for start, bits, name in fields:
        assert start >= pos
        if start > pos:
            channels.append(render.Channel(start - pos, 0, 0, 0))
        channels.append(render.Channel(bits, name == 'R', name == 'G', name == 'B', name == 'A'))
        pos = start + bits
    assert pos <= nbytes * 8
    return channels


class PygameRenderer(render.CallbackRenderer):
    """PygameRenderer([threads=0]) (lib/ntracer/pygame_render.py:51-117): draws the scene onto a pygame.Surface and, on
    completion, posts an event of type ON_COMPLETE with the attributes `source` (this renderer), `surface` and `scene`.
    The whole surface is drawn (clipping areas and subsurface boundaries are not honoured)."""

    #: event type sent when a frame is complete; any value between pygame.USEREVENT and pygame.NUMEVENTS may be assigned
    ON_COMPLETE = pygame.USEREVENT

    instances = weakref.WeakSet()

    def __init__(self, threads=0):
        super().__init__(threads)
        PygameRenderer.instances.add(self)
        self.last_channels = (None, None)        # (surface format the cached channel list was made for, the list)

    def begin_render(self, surface, scene):
        """Begin rendering `scene` onto `surface`; raises if the renderer is already running."""
        def done(_renderer):
            pygame.event.post(pygame.event.Event(self.ON_COMPLETE, source=self, scene=scene, surface=surface))

        layout = (surface.get_bitsize(), surface.get_masks())
        if layout != self.last_channels[0]:
            self.last_channels = (layout, channels_from_surface(surface))
        dest = surface.get_view() if hasattr(surface, 'get_view') else surface.get_buffer()
        fmt = render.ImageFormat(surface.get_width(), surface.get_height(), self.last_channels[1], surface.get_pitch(),
                                 pygame.get_sdl_byteorder() == pygame.LIL_ENDIAN)
        super().begin_render(dest, fmt, scene, done)


def _stop_renderers():
    # Pygame destroys its surfaces on shutdown whatever their reference counts: frames in flight must end first
    for r in list(PygameRenderer.instances):
        r.abort_render()


pygame.register_quit(_stop_renderers)
