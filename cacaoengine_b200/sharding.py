"""Work partitioning across the GPUs of one box (one process per GPU).

The hot path has no exchange step: pixels, images and cube rows are independent, so ranks take
disjoint shards and never talk to each other on the data path.  The only
cross-rank traffic is the barrier / max-reduction that brackets a timed region and the final
gather of results to the host.

  * conversion batches shard by image index           -> image_range()
  * a cubemap shards by (face, row band) work units    -> cube_units() / units_for_rank()
    and each rank needs only the source rows its units read -> source_window()
"""
from __future__ import annotations

from dataclasses import dataclass

from . import device

FACE_NAMES = ("right", "left", "up", "down", "front", "back")  # +X,-X,+Y,-Y,+Z,-Z (tools/cubetool/commands.hpp:31)


def image_range(rank: int, world: int, count: int) -> tuple[int, int]:
	"""Contiguous, balanced range of image indices [lo, hi) for `rank` (sizes differ by at most 1)."""
	if not (0 <= rank < world):
		raise ValueError(f"rank {rank} outside world of {world}")
	base, extra = divmod(count, world)
	lo = rank * base + min(rank, extra)
	return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class CubeUnit:
	face: int
	row_begin: int
	row_end: int

	@property
	def rows(self) -> int:
		return self.row_end - self.row_begin


def cube_units(face_size: int, bands_per_face: int, polar_factor: int = 1) -> list[CubeUnit]:
	"""Work units of a cube: bands_per_face row bands on each side face and polar_factor times as
	many on the +-Y faces (their rows cost up to twice as much, so they are cut finer), band-major so
	that consecutive units are different faces."""
	units = []
	for f in range(6):
		n = bands_per_face * (polar_factor if f in (2, 3) else 1)
		n = max(1, min(n, face_size))
		for b in range(n):
			r0 = face_size * b // n
			r1 = face_size * (b + 1) // n
			if r1 > r0:
				units.append(CubeUnit(f, r0, r1))
	return sorted(units, key=lambda u: (u.row_begin / face_size, u.face))


def bands_for_world(world: int) -> int:
	"""Bands per face: at least 8 units per rank's worth of granularity so the cost-aware dealing
	below can even out the polar faces (48 units for 8 ranks), never fewer than 1."""
	if world <= 1:
		return 1
	bands = 1
	while 6 * bands < 6 * world:
		bands *= 2
	return bands


def unit_cost(u: CubeUnit, face_size: int) -> float:
	"""Relative cost of a unit.  Measured on B200 (dev/unit_costs.py, 16384x8192 -> 4096^2, 8 bands per
	face: 15.2 us for every side-face band; 15.4 / 17.4 / 21.3 / 32.1 us from the edge to the centre
	of a +-Y face): a side-face row costs 1, a row of the +-Y faces 1.0 at the face edge rising to 2.1
	at the centre, where neighbouring pixels fan out over all longitudes and taps stop coalescing."""
	if u.face not in (2, 3):
		return float(u.rows)
	total = 0.0
	for r in range(u.row_begin, u.row_end):
		d = abs((2 * r + 1) / face_size - 1.0)  # 0 at the centre row, 1 at the edge
		total += 1.0 + 1.57 * (1.0 - d) ** 2.75
	return total


def units_for_rank(rank: int, world: int, face_size: int, bands_per_face: int | None = None) -> list[CubeUnit]:
	"""Cost-balanced dealing (longest-processing-time first): units sorted by estimated cost, each
	handed to the currently least-loaded rank.  Deterministic, so every rank computes the same plan
	without talking to the others."""
	if not (0 <= rank < world):
		raise ValueError(f"rank {rank} outside world of {world}")
	bands = bands_for_world(world) if bands_per_face is None else bands_per_face
	# the +-Y faces are cut twice as fine: their centre bands would otherwise be the largest units by
	# a factor of two and cap the balance (8 ranks: max/mean 1.04 -> 1.002)
	units = cube_units(face_size, bands, polar_factor=2 if world > 1 else 1)
	order = sorted(range(len(units)), key=lambda k: (-unit_cost(units[k], face_size), k))
	load = [0.0] * world
	mine = []
	for k in order:
		r = min(range(world), key=lambda q: (load[q], q))
		load[r] += unit_cost(units[k], face_size)
		if r == rank:
			mine.append(units[k])
	return sorted(mine, key=lambda u: (u.face, u.row_begin))


def source_window(units: list[CubeUnit], W: int, H: int, face_size: int) -> tuple[int, int]:
	"""(first_row, rows): the band of equirect rows that covers every unit in the list
	(host arithmetic identical to the kernel's, cacao_img_equirect_src_window)."""
	lo, hi = None, None
	for u in units:
		r0, n = device.equirect_src_window(W, H, face_size, 1 << u.face, u.row_begin, u.row_end)
		if n == 0:
			continue
		lo = r0 if lo is None else min(lo, r0)
		hi = r0 + n if hi is None else max(hi, r0 + n)
	return (0, 0) if lo is None else (lo, hi - lo)


def max_over_ranks(value: float) -> float:
	"""Multi-GPU timings are the max over ranks; single-process callers get the value back."""
	try:
		import torch
		import torch.distributed as dist
	except ImportError:
		return value
	if not (dist.is_available() and dist.is_initialized()):
		return value
	dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
	t = torch.tensor([value], dtype=torch.float64, device=dev)
	dist.all_reduce(t, op=dist.ReduceOp.MAX)
	return float(t.item())
