import sys,json; sys.path.insert(0,'/root/repo')
import bench
print(json.dumps(bench.train_16m_regime(bench.load_peaks())))
