"""Prints the switch list ("K1=V1,K2=V2", keys without the STARNEIG_B200_ prefix) of the fastest configuration in a
tools/sweep.py log that returned 0 and kept the Hessenberg form; prints nothing when that is the default configuration
or when it is less than 0.5 % faster than the default.   usage: best_of_sweep.py gpurun_out/sweep.log"""
import re, sys

rows = []
for line in open(sys.argv[1]):
    m = re.match(r"\[(.*?)\s*\] ret (-?\d+) device_ms\s+([0-9.]+).*form_ok (True|False)", line)
    if m and m.group(2) == "0" and m.group(4) == "True":
        rows.append((float(m.group(3)), m.group(1).strip()))
if rows:
    default = min((t for t, c in rows if c == "default"), default=None)
    t, cfg = min(rows)
    if cfg != "default" and (default is None or t < 0.995 * default):
        print(cfg)
