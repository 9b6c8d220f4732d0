#!/usr/bin/env python3
"""Turns the ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) into the per-kernel share
table kept under profiles/:  python tools/launch_shares.py X.csv "header line" ... > profiles/rNN_launch_shares.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    tot[name] = tot.get(name, 0) + int(r[-1])
    cnt[name] += 1
allt = sum(tot.values())
is_msm = lambda k: k.startswith(("msm::", "void msm::", "ba::", "void ba::"))  # noqa: E731  (ncu prints msm::ba:: kernels as ba::)
msm = sum(v for k, v in tot.items() if is_msm(k))
for h in sys.argv[2:]:
    print("# " + h)
print("\n| kernel | launches | total | share of all | share of the MSM |\n|---|---:|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("| %s | %d | %d | %.1f%% | %s |" % (k, cnt[k], v, 100 * v / allt, ("%.1f%%" % (100 * v / msm)) if is_msm(k) else ""))
