#!/usr/bin/env python
"""Print tools/ws_trace.py's JSON as a table: python tools/show_trace.py FILE..."""
import json
import sys
for f in sys.argv[1:]:
    d = json.load(open(f))
    print("==", f, "period", round(d["period_clk"], 1))
    for k, v in d["epilogue"]:
        print("E %-55s %8.1f" % (k, v))
    for k, v in d["issuer"]:
        print("I %-55s %s" % (k, v))
    for k, v in d["cross"].items():
        print("X %-60s %8.1f" % (k, v))
