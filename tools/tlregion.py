import json,sys
ev=json.load(open(sys.argv[1]))
lo,hi=float(sys.argv[2]),float(sys.argv[3])
for s,d,st,n in ev:
    if lo<=s<=hi and d>=2.5: print("%8.1f %7.1f  s%-4d %s"%(s,d,st,n[:60]))
