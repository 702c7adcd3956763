import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glob
from kafka_topic_analyzer_b200 import _native as N
variants = {
 'base': [],
 'stage1': ['KTA_EXP_ALIVE_STAGE=1'],
 'stage3': ['KTA_EXP_ALIVE_STAGE=3'],
}
for f in glob.glob(N.LIB_PATH.replace('.so','_exp_*.so')): os.remove(f)
import concurrent.futures as cf
def b(kv):
    k,v=kv
    N.build(force=True, defines=v, out=N.LIB_PATH.replace('.so','_exp_%s.so'%k)); return k
with cf.ThreadPoolExecutor(8) as ex:
    for k in ex.map(b, variants.items()): print('built',k)
