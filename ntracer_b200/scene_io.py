"""Flat scene files (.npz): the arrays of ntr_scene_desc plus camera."""
import numpy as np


def load_scene(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files if not k.startswith('g_')}


def save_scene(path, scene):
    np.savez_compressed(path, **{k: v for k, v in scene.items() if not k.startswith('_')})
