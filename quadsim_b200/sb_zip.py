"""Reader / writer for stable-baselines 2.x model archives.

Layout written by `BaseRLModel._save_to_file_zip` and used by the reference's shipped policies (`trained_model/best_model_v0.zip`,
`ppo2_docking*.zip`; saved through `PPO2.save`, rl_baselines/ppo2/ppo2.py:409-435): a zip with three members --
`data` (JSON of constructor arguments), `parameters` (an .npz of the TF variables), `parameter_list` (JSON list giving the
variable order).  For the docking MlpPolicy the variables are `model/{shared_fc0,pi_fc0,vf_fc0,vf,pi,q}/{w,b}:0` and
`model/pi/logstd:0` (weights stored [in, out]).
"""
import io
import json
import zipfile
from collections import OrderedDict

import numpy as np


def load_sb_zip(path):
    """-> (data: dict, params: OrderedDict[name -> np.ndarray]) in the archive's parameter order."""
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        data = json.loads(z.read("data").decode()) if "data" in names else {}
        npz = np.load(io.BytesIO(z.read("parameters")))
        order = json.loads(z.read("parameter_list").decode()) if "parameter_list" in names else list(npz.files)
        params = OrderedDict((k, np.asarray(npz[k])) for k in order)
    return data, params


def save_sb_zip(path, params, data=None):
    """Write `params` (name -> array) and JSON-serialisable `data` in the same three-member layout."""
    buf = io.BytesIO()
    np.savez(buf, **{k: np.asarray(v) for k, v in params.items()})
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(data or {}))
        z.writestr("parameters", buf.getvalue())
        z.writestr("parameter_list", json.dumps(list(params.keys())))
    return path
