import json,re,sys
txt=open(sys.argv[1]).read()
runs=re.findall(r"ensemble timeline \(ms since the first upload\): (\[.*\])", txt)
for r in runs:
    tl=json.loads(r)
    h=[round(m["h2d"][1]-m["h2d"][0]) for m in tl]
    c=sorted(set((m["compute"][0],m["compute"][1]) for m in tl))
    cd=[round(b-a) for a,b in c]
    gaps=[round(c[i+1][0]-c[i][1]) for i in range(len(c)-1)]
    hs=[round(m["h2d"][0]) for m in tl]
    print("end",round(tl[-1]["done"]),"h2d dur",h,"h2d start",hs,"compute",cd,"gaps",gaps)
