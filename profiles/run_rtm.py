import sys,json; sys.path.insert(0,'/root/repo')
import bench
print(json.dumps(bench.rtm_regime(bench.load_peaks())))
