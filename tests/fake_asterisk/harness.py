"""ctypes driver of tests/_build/app_tiresias_dropin.so (TEST INFRASTRUCTURE): plays the Asterisk core for the
reference's unchanged module shell, dialplan application and CLI running on the replacement fp_handler.c."""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

from . import build as _build


class FakeAsterisk:
    def __init__(self, root: str, backup_db: str | None = None, device: int = 0):
        so = _build.build()
        if so is None:
            raise RuntimeError("the drop-in module is not built and /root/reference is absent")
        os.makedirs(os.path.join(root, "etc", "asterisk"), exist_ok=True)
        os.makedirs(os.path.join(root, "var", "lib", "asterisk", "third-party"), exist_ok=True)
        os.environ["FAKE_AST_ROOT"] = root
        os.environ["TIRESIAS_BACKUP_DATABASE"] = backup_db or os.path.join(root, "var", "lib", "asterisk", "third-party", "tiresias", "audio_recongition.db")
        os.environ["TIRESIAS_GPU_DEVICE"] = str(device)
        self.root = root
        L = self.L = C.CDLL(so)
        L.fake_module_load.restype = C.c_int
        L.fake_module_unload.restype = C.c_int
        L.fake_module_reload.restype = C.c_int
        L.fake_module_description.restype = C.c_char_p
        L.fake_cli_run.argtypes = [C.c_char_p, C.c_int]
        L.fake_cli_count.restype = C.c_int
        L.fake_take_log.restype = C.c_void_p
        L.fake_free.argtypes = [C.c_void_p]
        L.fake_channel_new.restype = C.c_void_p
        L.fake_channel_new.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_long, C.c_int]
        L.fake_channel_free.argtypes = [C.c_void_p]
        L.fake_channel_var.restype = C.c_char_p
        L.fake_channel_var.argtypes = [C.c_void_p, C.c_char_p]
        L.fake_channel_answered.argtypes = [C.c_void_p]
        L.fake_channel_samples_read.restype = C.c_long
        L.fake_channel_samples_read.argtypes = [C.c_void_p]
        L.fake_app_exec.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p]

    def write_conf(self, text: str, name: str = "tiresias.conf"):
        with open(os.path.join(self.root, "etc", "asterisk", name), "w") as f:
            f.write(text)

    def load(self) -> int:
        return self.L.fake_module_load()

    def unload(self) -> int:
        return self.L.fake_module_unload()

    def log(self) -> str:
        p = self.L.fake_take_log()
        s = C.string_at(p).decode(errors="replace")
        self.L.fake_free(p)
        return s

    def cli(self, line: str):
        """-> (rc, output text); rc 0 success, 1 usage, 2 failure, -1 unknown command"""
        with tempfile.TemporaryFile() as f:
            rc = self.L.fake_cli_run(line.encode(), f.fileno())
            f.seek(0)
            return rc, f.read().decode()

    def exec_app(self, data: str, pcm: np.ndarray, rate: int = 8000, state_up: bool = False, hangup_at: int = -1, control_every: int = 0,
                 app: str = "Tiresias"):
        """run the dialplan application on a scripted channel -> (rc, {channel variables}, info)"""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        ch = self.L.fake_channel_new(pcm.ctypes.data, pcm.size, rate, int(state_up), hangup_at, control_every)
        try:
            rc = self.L.fake_app_exec(app.encode(), ch, data.encode())
            names = ["TIRSTATUS", "TIRFRAMECOUNT", "TIRMATCHCOUNT", "TIRFILEUUID", "TIRFILENAME", "TIRCONTEXT", "TIRFILEHASH"]
            out = {}
            for n in names:
                v = self.L.fake_channel_var(ch, n.encode())
                if v is not None:
                    out[n] = v.decode()
            info = {"answered": self.L.fake_channel_answered(ch), "samples_read": self.L.fake_channel_samples_read(ch)}
            return rc, out, info
        finally:
            self.L.fake_channel_free(ch)
