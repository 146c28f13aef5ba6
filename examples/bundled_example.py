"""The reference's documented workflow (README.md:21-25; man/*.Rd examples) on the bundled basin,
through the Python mirror of the R API.  Needs a CUDA device and the built library.

Inputs are the decoded fixtures of tests/golden/c1_inputs.npz (raw/roi.*, dem.tif, pisco_*.tif and
qobs.txt of the reference, decoded by tests/golden/make_fixtures.py), so it also runs where
/root/reference is not available.
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as g  # noqa: E402
from gr2msemidistr_b200 import api, gis  # noqa: E402

g.build()
inp = np.load(os.path.join(ROOT, "tests", "golden", "c1_inputs.npz"))
n = len(inp["area"])
cols = {"DatesR": [f"{1981 + t // 12:04d}-{t % 12 + 1:02d}-01" for t in range(432)]}
cols.update({f"P_{c}": inp["P"][i] for i, c in enumerate(inp["comid"])})
cols.update({f"E_{c}": inp["E"][i] for i, c in enumerate(inp["comid"])})
cols["Q"] = inp["qobs"]
data = pd.DataFrame(cols)
xy = inp["poly_xy"]
roi = gis.Subbasins(inp["area"], np.array(["A"] * n), inp["comid"], [[xy[xy[:, 0] == i][:, 1:]] for i in range(n)])
dem = gis.Raster(np.where(inp["dem"] == -32768, -3e38, inp["dem"]).astype(np.float32)[None], *inp["dem_geo"], nodata=-3e38)

# Run_GR2MSemiDistr(Data=data, Subbasins=roi, RunIni='01/1981', RunEnd='12/2016', Parameters=c(10.976, 0.665, 1.186, 1.169))
model = api.Run_GR2MSemiDistr(data, roi, "01/1981", "12/2016", Parameters=[10.976, 0.665, 1.186, 1.169])
print("outlet discharge, first year [m3/s]: sim", model["SINK"]["sim"][:12], "obs", model["SINK"]["obs"][:12])

# Routing_GR2MSemiDistr(Model=model, Subbasins=roi, Dem=dem, AcumIni='01/1981', AcumEnd='12/2016')
routed = api.Routing_GR2MSemiDistr(model, roi, dem, AcumIni="01/1981", AcumEnd="12/2016")
print("routed discharge at each subbasin, 01/1981:", routed["QR"][0])

# Optim_GR2MSemiDistr(..., RunIni='01/1981', RunEnd='12/2002', WarmUp=36, Parameters=c(1000,1,1,1), ... Max.Functions=1000, 'KGE')
cal = api.Optim_GR2MSemiDistr(data, roi, "01/1981", "12/2002", WarmUp=36, Parameters=[1000, 1, 1, 1],
                              Parameters_Min=[1, 0.01, 0.8, 0.8], Parameters_Max=[2000, 2, 1.2, 1.2],
                              Max_Functions=1000, Optimization="KGE", restarts=64)
print("calibrated:", cal["Param"], "KGE", cal["Value"], "in", round(cal["elapsed_s"], 2), "s,", cal["evaluations"], "evaluations")
