#!/usr/bin/env python
"""The metadata pass (vocabulary + edge types, ``model.compute_metadata``) over shard files: raw datapoints through the
host-language chain vs ``ShardDataset.update_model_metadata`` (native decoder, counts per worker thread).  Host cores only.

    python scripts/bench_metadata.py [directory with *.msgpack.l.gz; default: 4 synthetic 500-graph shards]
"""
import os, sys, time, tempfile, json
ROOT="/root/repo"; PKG=os.path.join(ROOT,"neurips21-self-supervised-bug-detection-and-repair_b200")
sys.path[:0]=[ROOT,PKG]
from pathlib import Path
from buglab.models.modelregistry import load_model
from buglab.utils.msgpackutils import load_all_msgpack_l_gz
from buglab_b200.shards import ShardDataset
from dpu_utils.utils import RichPath
from buglab_b200.synthetic import write_shards
if len(sys.argv) > 1:
    d = sys.argv[1]
else:
    d = tempfile.mkdtemp(prefix="buglab_metadata_")
    write_shards(d, 4, 500, seed=11)
rich=RichPath.create(d)
total = sum(1 for _ in load_all_msgpack_l_gz(rich))
def fresh():
    m,_,_=load_model({"modelName":"gnn-mlp","hidden_state_size":16}, Path("/tmp/_m.pkl.gz")); return m
res={}
a=fresh(); t=time.perf_counter(); a.compute_metadata(load_all_msgpack_l_gz(rich)); res["host_graphs_per_s"]=round(total/(time.perf_counter()-t),1)
for th in (1,2,4,8):
    b=fresh(); t=time.perf_counter(); b.compute_metadata(ShardDataset(rich,num_threads=th)); res[f"native_{th}thread_graphs_per_s"]=round(total/(time.perf_counter()-t),1)
    assert a.gnn_model.node_representation_model.vocabulary.token_to_id==b.gnn_model.node_representation_model.vocabulary.token_to_id
print(json.dumps({"metric":"metadata pass (vocabulary + edge types), graphs/s", "graphs": total,"cores":os.cpu_count(),"results":res}))
