"""Hot spots of a kernel from an .ncu-rep captured with --import-source on: instructions executed and stall samples per
source line (ncu -i X --page source --csv, SASS view).  Run here, no GPU.
    python tools/ncu_source_hot.py gpurun_out/ncu_r02_k_nn.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    isrc = H.index("Source")
    iins = H.index("Instructions Executed")
    ismp = H.index("# Samples") if "# Samples" in H else None
    data = []
    for r in rows[hdr + 1:]:
        if len(r) <= iins:
            continue
        try:
            data.append((int(r[iins]), int(r[ismp]) if ismp is not None and r[ismp] else 0, r[isrc].strip()))
        except ValueError:
            pass
    tot_i = sum(d[0] for d in data) or 1
    tot_s = sum(d[1] for d in data) or 1
    print(f"rows {len(data)} instructions {tot_i} samples {tot_s}")
    print("--- by instructions executed")
    for ins, smp, src in sorted(data, key=lambda d: -d[0])[:top]:
        print(f"{100 * ins / tot_i:5.1f}% inst {100 * smp / tot_s:5.1f}% smp  {src[:110]}")
    print("--- by stall samples")
    for ins, smp, src in sorted(data, key=lambda d: -d[1])[:top]:
        print(f"{100 * ins / tot_i:5.1f}% inst {100 * smp / tot_s:5.1f}% smp  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
