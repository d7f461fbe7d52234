"""Extract the 69-channel ERA5 statistics the DA driver hard-codes.

Reads (never imports) /root/reference/da_4dvar.py and pulls out three numeric
lists by their position in the source:
  * mean_layer / std_layer   da_4dvar.py:641-643 (get_model_mean_std)
  * stdTr                    da_4dvar.py:1181    (vae4dvar background-error scaling)
and writes them to vaevar_b200/data/era5_stats.json.  The JSON is data, not code;
it is committed so nothing at run time needs /root/reference.
"""
import json, re, sys, pathlib

src = pathlib.Path("/root/reference/da_4dvar.py").read_text()


def grab(pattern):
    m = re.search(pattern, src, re.S)
    if not m:
        sys.exit("pattern not found: " + pattern)
    return [float(t) for t in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", m.group(1))]


mean = grab(r"mean_layer = np\.array\(\[(.*?)\]\)")
std = grab(r"std_layer = np\.array\(\[(.*?)\]\)")
stdtr = grab(r'mode == "vae4dvar":\s*stdTr = torch\.Tensor\(\[(.*?)\]\)')
assert len(mean) == len(std) == len(stdtr) == 69, (len(mean), len(std), len(stdtr))
out = pathlib.Path(__file__).resolve().parents[1] / "vaevar_b200" / "data" / "era5_stats.json"
out.write_text(json.dumps({
    "source": "da_4dvar.py:641-643 (mean,std), da_4dvar.py:1181 (stdTr)",
    "mean": mean, "std": std, "stdTr": stdtr}, indent=0))
print("wrote", out)
