"""Host-side layout maths for multi-GPU rendering.  Pure index arithmetic, no compute.

Tile mode: the frame is cut into bands of `band_rows` rows; band b belongs to rank b % world and is that rank's local
band b // world.  Every rank renders `bands_per_rank * band_rows` local rows (tail bands past the image are padding),
one all-gather of equal-sized blocks assembles the frame.  The mapping matches rtb_trace_args (include/rtb200.h):
    y(j) = ((j // band_rows) * world + rank) * band_rows + j % band_rows
Sample-range mode: rank r renders samples [r * spp / world, (r + 1) * spp / world) of every pixel (sampleSkip = first).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class BandLayout:
    height: int
    world: int
    band_rows: int

    @property
    def bands(self) -> int:
        return (self.height + self.band_rows - 1) // self.band_rows

    @property
    def bands_per_rank(self) -> int:
        return (self.bands + self.world - 1) // self.world

    @property
    def local_rows(self) -> int:
        return self.bands_per_rank * self.band_rows

    def global_row(self, rank: int, j: int) -> int:
        """global image row of local row j of `rank` (may be >= height: padding)"""
        return ((j // self.band_rows) * self.world + rank) * self.band_rows + j % self.band_rows

    def owned_rows(self, rank: int) -> list[int]:
        return [y for y in (self.global_row(rank, j) for j in range(self.local_rows)) if y < self.height]


def single_gpu_layout(height: int) -> BandLayout:
    return BandLayout(height, 1, height)


def assemble_gathered(gathered, layout: BandLayout):
    """gathered: [world, local_rows, W, C] (torch tensor or numpy array) -> [height, W, C]"""
    w, b = layout.world, layout.band_rows
    tail = tuple(gathered.shape[2:])
    g = gathered.reshape(w, layout.bands_per_rank, b, *tail)
    if hasattr(g, "permute"):
        g = g.permute(1, 0, 2, *range(3, g.dim()))
    else:
        g = g.transpose(1, 0, 2, *range(3, g.ndim))
    return g.reshape(layout.bands_per_rank * w * b, *tail)[: layout.height]


def sample_range(spp: int, world: int, rank: int) -> tuple[int, int]:
    """(first sample, sample count) of `rank` in sample-range mode"""
    lo = (spp * rank) // world
    hi = (spp * (rank + 1)) // world
    return lo, hi - lo
