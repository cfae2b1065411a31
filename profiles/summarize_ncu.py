"""Turn an `ncu --set full` report into the compact per-launch summary kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/r1_full.ncu-rep profiles/r1_full_summary.csv [--traffic profiles/traffic.json]

Columns: the metrics the roofline argument rests on (duration, DRAM bytes, DRAM/L2/L1 throughput,
issue activity, occupancy, registers, top stall reasons).  --traffic also writes DRAM bytes per
launch (read + write) keyed by the names bench.py looks up.
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
	"gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
	"lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
	"lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
	"launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
	"sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
	"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
	"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
# kernel-name fragment -> key used by bench.py traffic_for()
TRAFFIC_KEYS = {
	"convert_tiles_kernel<0, 2, 4, 4": "cfg2_rgba8_to_rgba16f",
	"convert_tiles_kernel<0, 0, 3, 4": "cfg4a_16x_4096sq_rgb8_to_rgba8_per_launch",
	"convert_tiles_kernel<0, 3, 3, 4": "cfg4b_16x_4096sq_rgb8_to_rgba32f_per_launch",
	"convert_tiles_kernel<1, 0, 4, 3": "lut_64x_2048sq_rgba16_to_rgb8",
	"convert_wide_kernel<1, 0, 4, 3, 0, 0": "lut_64x_2048sq_rgba16_to_rgb8",  # plain (not the FLIP instantiation)
	"convert_wide_kernel<1, 0, 4, 3, 0>": "lut_64x_2048sq_rgba16_to_rgb8",    # round-1 name, before FLIP was a template argument
	"convert_lanes_kernel<4, 3, 0": "lut_64x_2048sq_rgba16_to_rgb8",                 # bank-private table form (convert_lanes.cuh), plain
	"convert_lanes_kernel<4, 3, 1": "engine_cube_6x2048sq_rgba16_to_rgb8_flipped",   # ... FLIP instantiation
	"equirect_kernel<3, 4": "cfg3_equirect_8192x4096_rgba32f_to_6x2048",
	"convert_wide_kernel<1, 0, 4, 3, 0, 1": "engine_cube_6x2048sq_rgba16_to_rgb8_flipped",  # FLIP instantiation
	"equirect_rows_kernel<0, 4": "cfg3shape_equirect_8192x4096_rgba8_to_6x2048",
	"equirect_rows_kernel<0, 1": "cfg3shape_equirect_8192x4096_gray8_to_6x2048",
	"equirect_rows_kernel<0, 3": "cfg3shape_equirect_8192x4096_rgb8_to_6x2048",
	"equirect_kernel<1, 4": "cfg3shape_equirect_8192x4096_rgba16_to_6x2048",
	"equirect_kernel<2, 4": "cfg3shape_equirect_8192x4096_rgba16f_to_6x2048",
}
# equirect launches are told apart by their grid (same kernel for cfg3 and cfg5)
GRID_KEYS = {
	("equirect_kernel<3, 4", "(128, 64, 6)"): "cfg3_equirect_8192x4096_rgba32f_to_6x2048",
	("equirect_kernel<3, 4", "(256, 128, 6)"): "cfg5_equirect_16384x8192_rgba32f_to_6x4096",
	("equirect_kernel<0, 4", "(32, 16, 6)"): "cfg1_equirect_2048x1024_rgba8_to_6x512",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
	rep, out = sys.argv[1], sys.argv[2]
	raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
	rows = list(csv.reader(io.StringIO(raw)))
	hdr, units = rows[0], rows[1]
	idx = {h: i for i, h in enumerate(hdr)}
	traffic = {}
	with open(out, "w", newline="") as f:
		w = csv.writer(f)
		w.writerow(["kernel"] + [f"{m} [{units[idx[m]]}]" for m in METRICS if m in idx])
		for r in rows[2:]:
			name = r[idx["Kernel Name"]]
			w.writerow([name[:90]] + [r[idx[m]] for m in METRICS if m in idx])
			grid = r[idx["Grid Size"]].replace(" ", "").replace(",", ", ") if "Grid Size" in idx else ""
			keys = dict(TRAFFIC_KEYS)
			for (frag, g), key in GRID_KEYS.items():
				if frag in name:
					keys.pop(frag, None)
					if g == grid:
						keys[frag] = key
			# the fused resample + convert instantiations carry the destination format / channels as trailing
			# template arguments: RGBA32F -> RGBA16F is equirect_kernel<3, 4, unsigned int, 1, 8, 2, 4>
			if "equirect_kernel<2, 4" in name and not name.split("equirect_kernel<2, 4")[1].startswith(", unsigned int, 1, 8, 2, 4>"):
				keys.pop("equirect_kernel<2, 4", None)  # a fused F16 instantiation, not the plain RGBA16F resampling
			if "equirect_kernel<3, 4" in name and ", 2, 4>(" in name:
				keys = {"equirect_kernel<3, 4": "cfg3_fused_rgba32f_to_rgba16f_cube"} if grid == "(128, 64, 6)" else {}
			for frag, key in keys.items():
				if frag in name:
					rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT_SCALE[units[idx["dram__bytes_read.sum"]]]
					wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT_SCALE[units[idx["dram__bytes_write.sum"]]]
					traffic[key] = int(rd + wr)  # last launch of that kernel wins (warm)
	if "--traffic" in sys.argv:
		p = sys.argv[sys.argv.index("--traffic") + 1]
		old = {}
		try:
			old = json.load(open(p))
		except Exception:
			pass
		old.update(traffic)
		json.dump(old, open(p, "w"), indent=1, sort_keys=True)
	print("wrote", out, "traffic:", traffic)


if __name__ == "__main__":
	main()
