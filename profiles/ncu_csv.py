"""Print (kernel, metric, value) rows of an `ncu --csv --log-file` launch list."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
    print(f"{name:32s} {row['Metric Name']:44s} {row['Metric Value']:>14s} {row['Metric Unit']}")
