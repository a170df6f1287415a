"""Developer tool: aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list over the LAST training step
(one stem-to-stem period at the end of the list) into a per-kernel table (markdown)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if l.startswith('"')]))
    idx = [i for i, x in enumerate(rows) if "stem_im2col" in x["Kernel Name"]]
    # one step = the launches between two consecutive stem kernels (BERT is enqueued ahead of the stem on its own lane)
    step = rows[-(idx[-1] - idx[-2]):] if len(idx) >= 2 else rows
    tot = sum(float(x["Metric Value"]) for x in step) / 1e3
    agg = collections.OrderedDict()
    for x in step:
        n = re.sub(r"\(.*", "", x["Kernel Name"].replace("(int)", ""))[:70]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(x["Metric Value"]) / 1e3
    print(f"one step: {len(step)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised)\n")
    print("| kernel | launches | us | us/launch | share |\n|---|---|---|---|---|")
    for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if us / tot < 0.001:
            continue
        print(f"| `{k}` | {c} | {us:.0f} | {us / c:.1f} | {100 * us / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
