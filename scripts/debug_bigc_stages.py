"""Stage-by-stage comparison of the device BIG-C forward with the oracle (debug aid; runs on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from vidsgg_big_b200 import synth, bigc
from oracle import bigc as ob
from oracle.geometry import stretch_index_map

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32_simt"
cfg = synth.tiny_vidvrd_config()
st = synth.make_bigc_state(7, cfg)
P = synth.make_proposal(202, 12, 64, cfg["dim_feat"] + cfg["dim_i3d"], cfg["num_enti_cats"], max_len=30)
E = cfg["dim_enti"]

def rel(a, b):
    a = a.detach().cpu().double().numpy(); b = b.detach().cpu().double().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-9))

# ---- oracle intermediates on unique frames
w, h = P.video_wh
bl, fl = P.bboxes_list, P.features_list
Tmax = max(b.shape[0] for b in bl)
f8 = torch.cat([ob.box_motion_features(b, w, h) for b in bl], 0)
h1 = F.relu(F.linear(f8, st["fc_bbox2enti.0.weight"], st["fc_bbox2enti.0.bias"]))
xb = F.relu(F.linear(h1, st["fc_bbox2enti.2.weight"], st["fc_bbox2enti.2.bias"]))
feats = torch.cat(fl, 0)
xv = F.relu(F.linear(F.relu(F.linear(feats[:, :cfg["dim_feat"]], st["fc_feat2enti.0.weight"], st["fc_feat2enti.0.bias"])), st["fc_feat2enti.2.weight"], st["fc_feat2enti.2.bias"]))
X = torch.cat([xb, xv], 1)
cw = st["conv_feat2enti.weight"]
Y = torch.cat([X @ cw[:, :, k].t() for k in range(3)], 1)
q, logits, att, inter = ob.encode2decode(st, cfg, P, return_intermediates=True)
# oracle encoder / decoder per layer
enc = []
x = inter["enti2enco"]
for i in range(cfg["n_enco_layers"]):
    x = ob.encoder_layer(st, "encoder_layers.%d" % i, x, cfg["n_att_head"]); enc.append(x)
dec = []
qq = st["pred_query_init"]
for i in range(cfg["n_deco_layers"]):
    qq, a_ = ob.decoder_layer(st, "decoder_layers.%d" % i, qq, st["pos_embedding"], x, cfg["n_att_head"], cfg["dim_att"], E); dec.append(qq)

model = bigc.BIG_C_vidvrd(cfg, precision=prec); model.load_state_dict(st); model.cuda()
model._dbg = {}
P.to("cuda:0")
with torch.no_grad():
    query, lg, at, so, ex = model.forward_debug(P)
d = model._dbg
print("bbox_h1   ", rel(d["bbox_h1"], h1))
print("X bbox    ", rel(d["X"][:, :E], xb))
print("X feat    ", rel(d["X"][:, E:], xv))
print("Y taps    ", rel(d["Y"], Y))
print("pooled    ", rel(d["pooled"], inter["pooled"]))
print("enti2enco ", rel(ex["enti2enco"], inter["enti2enco"]))
print("extra     ", rel(ex["extra"], inter["extra_avg"]))
for i, (a, b) in enumerate(zip(d["enc_layers"], enc)): print("enc layer", i, rel(a, b))
for i, (a, b) in enumerate(zip(d["dec_layers"], dec)): print("dec layer", i, rel(a, b))
print("att       ", rel(at, att))
print("logits    ", rel(lg, logits))
