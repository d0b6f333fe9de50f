import glob, json, sys
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-46s N=%d nb=%d grid=%s %s TF %.1f ll=%.6f ph=%s" % (f.split("/")[-1], d["N"], d["nb"], d["grid"], ["%.3f" % t for t in d["seconds"]], d["tflops_equiv"], d["ll"], {k: round(v, 1) for k, v in d["info"]["phases_ms"].items()}))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
