"""Golden-case table shared by tests/golden/make_golden.py (which runs the reference) and the
parity tests (which only read the committed .npz fixtures).  No reference import here."""
import numpy as np

CASES = {
    # name: (ctor kwargs, B, target_dims, seed)
    "tiny_v2": (dict(n_features=5, window_size=12, out_dim=5, kernel_size=3, gru_hid_dim=8,
                     forecast_n_layers=2, forecast_hid_dim=6, recon_hid_dim=7), 3, None, 1),
    "tiny_v1": (dict(n_features=5, window_size=12, out_dim=5, kernel_size=5, use_gatv2=False,
                     feat_gat_embed_dim=4, time_gat_embed_dim=3, gru_hid_dim=8,
                     forecast_n_layers=1, forecast_hid_dim=6, recon_hid_dim=7), 3, None, 2),
    "tiny_v2_embed_out1": (dict(n_features=6, window_size=10, out_dim=1, kernel_size=7,
                                feat_gat_embed_dim=4, time_gat_embed_dim=3, gru_hid_dim=9,
                                forecast_n_layers=3, forecast_hid_dim=5, recon_hid_dim=11), 4, [0], 3),
    "tiny_v1_default_embed": (dict(n_features=4, window_size=9, out_dim=4, kernel_size=3, use_gatv2=False,
                                   gru_hid_dim=6, forecast_n_layers=1, forecast_hid_dim=6,
                                   recon_hid_dim=6), 2, None, 4),
    "tiny_multilayer": (dict(n_features=5, window_size=12, out_dim=5, kernel_size=3, gru_n_layers=2,
                             gru_hid_dim=8, forecast_n_layers=2, forecast_hid_dim=6, recon_n_layers=2,
                             recon_hid_dim=7), 3, None, 5),
    "c1": (dict(n_features=25, window_size=100, out_dim=25), 4, None, 6),
    "c2_b8": (dict(n_features=38, window_size=100, out_dim=38, forecast_n_layers=3, dropout=0.3), 8, None, 7),
}


def inputs_for(cfg, B, seed):
    rng = np.random.default_rng(1000 + seed)
    x = rng.random((B, cfg.n, cfg.k))
    y = rng.random((B, 1, cfg.k))
    return x, y


