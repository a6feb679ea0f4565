"""Reference-named geometry helpers (models/common/ops.py) backed by the closed forms of csrc/grl_geometry.h."""
from ...geometry import coords_table, position_index, shift_mask  # noqa: F401


def get_relative_coords_table_all(window_size, pretrained_window_size=[0, 0], anchor_window_down_factor=1):
    if any(pretrained_window_size):
        raise NotImplementedError("pretrained window sizes are unused by every GRL config")
    return coords_table(tuple(window_size), anchor_window_down_factor)


def get_relative_position_index_simple(window_size, anchor_window_down_factor=1, window_to_anchor=True):
    return position_index(tuple(window_size), anchor_window_down_factor, window_to_anchor)


def calculate_mask(input_resolution, window_size, shift_size):
    return shift_mask(tuple(input_resolution), tuple(window_size), shift_size)


def calculate_mask_all(input_resolution, window_size, shift_size, anchor_window_down_factor=1, window_to_anchor=True):
    return shift_mask(tuple(input_resolution), tuple(window_size), shift_size, anchor_window_down_factor, window_to_anchor)
