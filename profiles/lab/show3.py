import json,sys
txt=open(sys.argv[1]).read().strip().split("\n")[-1]
d=json.loads(txt)
c=d["configs"]["3"]
print(sys.argv[1], "cfg2", round(d["ms_per_step"],3), "cfg3", round(c["ms_per_step"],3), "nocoll", c["collective"] and round(c["collective"]["ms_per_step_without_collective"],3), "exposed", c["collective"] and round(c["collective"]["ms_exposed"],3))
