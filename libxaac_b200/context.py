"""Context: one per process/GPU. Owns the device copy of the ROM tables and the host-API staging."""
import ctypes

from . import _lib


class Context:
    def __init__(self, device=0, imdct_rom=None, qmf_rom=None, env_rom=None, misc_rom=None, ps_rom=None, usac_rom=None, esbr_rom=None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        rc = self._lib.xaac_b200_create(ctypes.byref(self._h), int(device))
        if rc != 0:
            self._h = ctypes.c_void_p()
            raise _lib.XaacB200Error(
                f"xaac_b200_create(device={device}) failed with 0x{rc & 0xffffffff:08x}: "
                "a CUDA device is required (no CPU fallback)"
            )
        self.device = int(device)
        self.set_imdct_rom(imdct_rom if imdct_rom is not None else _lib.rom_blob("imdct_rom.bin"))
        self.set_qmf_rom(qmf_rom if qmf_rom is not None else _lib.rom_blob("qmf_rom.bin"))
        self.set_env_rom(env_rom if env_rom is not None else _lib.rom_blob("env_rom.bin"),
                         misc_rom if misc_rom is not None else _lib.rom_blob("misc_rom.bin"))
        self.set_ps_rom(ps_rom if ps_rom is not None else _lib.rom_blob("ps_rom.bin"))
        self.set_usac_rom(usac_rom if usac_rom is not None else _lib.rom_blob("usac_rom.bin"))
        self.set_esbr_rom(esbr_rom if esbr_rom is not None else _lib.rom_blob("esbr_rom.bin"))
        self.set_esbr_envcalc_rom(_lib.rom_blob("esbr_random_phase.bin"))
        self.set_hbe_rom(_lib.rom_blob("hbe_rom.bin"))
        self.set_fps_rom(_lib.rom_blob("fps_rom.bin"))
        self.set_block_rom(_lib.rom_blob("block_rom.bin"))

    # -- plumbing ---------------------------------------------------------------------------------
    @property
    def handle(self):
        return self._h

    def check(self, rc, what):
        if rc != 0:
            msg = self._lib.xaac_b200_last_error(self._h)
            raise _lib.XaacB200Error(f"{what} failed (0x{rc & 0xffffffff:08x}): {msg.decode() if msg else ''}")

    def set_imdct_rom(self, blob):
        """blob: bytes of the host's ia_aac_dec_imdct_tables_struct (>= 7500 leading bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_imdct_rom(self._h, buf, len(blob)), "xaac_b200_set_imdct_rom")

    def set_qmf_rom(self, blob):
        """blob: bytes of the host's ia_qmf_dec_tables_struct (>= 3464 leading bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_qmf_rom(self._h, buf, len(blob)), "xaac_b200_set_qmf_rom")

    def set_env_rom(self, env_blob, misc_blob):
        """env_blob: bytes of the host's ia_env_calc_tables_struct (2404); misc_blob: leading >= 2470 bytes of the
        host's ixheaacd_misc_tables."""
        e = (ctypes.c_char * len(env_blob)).from_buffer_copy(env_blob)
        m = (ctypes.c_char * len(misc_blob)).from_buffer_copy(misc_blob)
        self.check(self._lib.xaac_b200_set_env_rom(self._h, e, len(env_blob), m, len(misc_blob)),
                   "xaac_b200_set_env_rom")

    def set_ps_rom(self, blob):
        """blob: leading >= 1230 bytes of the host's ia_ps_tables_struct."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_ps_rom(self._h, buf, len(blob)), "xaac_b200_set_ps_rom")

    def set_usac_rom(self, blob):
        """blob: the host's USAC FD tables concatenated in the XAAC_UROM_* order (15880 bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_usac_rom(self._h, buf, len(blob)), "xaac_b200_set_usac_rom")

    def set_esbr_rom(self, blob):
        """blob: esbr_qmf_c | esbr_w_32 | esbr_sin_cos_twiddle_l64 | esbr_alt_sin_twiddle_l64 of the host's
        ia_qmf_dec_tables_struct (5744 bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_esbr_rom(self._h, buf, len(blob)), "xaac_b200_set_esbr_rom")

    def set_hbe_rom(self, blob):
        """The harmonic transposer's float tables (XAAC_HROM_* layout, 37296 bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_hbe_rom(self._h, buf, len(blob)), "xaac_b200_set_hbe_rom")

    def set_fps_rom(self, blob):
        """The float parametric stereo's tables (XAAC_FPSROM_* layout, 4064 bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_fps_rom(self._h, buf, len(blob)), "xaac_b200_set_fps_rom")

    def set_block_rom(self, blob):
        """The leading 620 bytes of ia_aac_dec_block_tables_struct (scale tables, TNS coefficient tables) for the spectral stage."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_block_rom(self._h, buf, len(blob)), "xaac_b200_set_block_rom")

    def set_esbr_envcalc_rom(self, blob):
        """blob: ixheaac_random_phase[512][2] (4096 bytes)."""
        buf = (ctypes.c_char * len(blob)).from_buffer_copy(blob)
        self.check(self._lib.xaac_b200_set_esbr_envcalc_rom(self._h, buf, len(blob)), "xaac_b200_set_esbr_envcalc_rom")

    @property
    def num_sms(self):
        return self._lib.xaac_b200_num_sms(self._h)

    @property
    def launch_count(self):
        return self._lib.xaac_b200_launch_count(self._h)

    def kernel_timing(self, enable):
        """bracket every kernel launch with CUDA events (profiling hook); clears previous records"""
        self.check(self._lib.xaac_b200_kernel_timing(self._h, int(bool(enable))), "xaac_b200_kernel_timing")

    def kernel_times(self):
        """{kernel name: (total ms, launches)} since kernel_timing(True)"""
        buf = ctypes.create_string_buffer(8192)
        self.check(self._lib.xaac_b200_kernel_times(self._h, buf, len(buf)), "xaac_b200_kernel_times")
        out = {}
        for rec in buf.value.decode().split(";"):
            if rec:
                name, ms, cnt = rec.split(":")
                out[name] = (float(ms), int(cnt))
        return out

    def sync(self):
        self.check(self._lib.xaac_b200_sync(self._h), "xaac_b200_sync")

    # ---- raw device buffers and peer (CUDA IPC) buffers: the multi-GPU I/O path, one process per GPU ----
    def dev_tensor(self, nbytes, dtype, shape):
        """A device buffer from xaac_b200_dev_alloc (a cudaMalloc base pointer, hence exportable) viewed as a torch tensor."""
        p = ctypes.c_void_p()
        self.check(self._lib.xaac_b200_dev_alloc(self._h, int(nbytes), ctypes.byref(p)), "xaac_b200_dev_alloc")
        return _RawBuffer(self, p.value, int(nbytes), owned=True).as_tensor(dtype, shape)

    def ipc_export(self, tensor):
        """64-byte CUDA IPC handle of a tensor made by dev_tensor()."""
        h = (ctypes.c_char * 64)()
        self.check(self._lib.xaac_b200_ipc_export(self._h, ctypes.c_void_p(tensor.data_ptr()), h), "xaac_b200_ipc_export")
        return bytes(h)

    def ipc_import(self, handle, nbytes, dtype, shape):
        """Map another rank's exported buffer on this context's device (peer access over NVLink) and view it as a tensor; the
        _dev entry points accept it like any device pointer."""
        p = ctypes.c_void_p()
        hb = (ctypes.c_char * 64).from_buffer_copy(handle)
        self.check(self._lib.xaac_b200_ipc_import(self._h, hb, ctypes.byref(p)), "xaac_b200_ipc_import")
        return _RawBuffer(self, p.value, int(nbytes), owned=False).as_tensor(dtype, shape)

    def close(self):
        if self._h:
            self._lib.xaac_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _RawBuffer:
    """Holder that exposes a raw device pointer through __cuda_array_interface__ (torch.as_tensor aliases it)."""

    def __init__(self, ctx, ptr, nbytes, owned):
        self.ctx, self.ptr, self.nbytes, self.owned = ctx, ptr, nbytes, owned
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}

    def as_tensor(self, dtype, shape):
        import torch
        t = torch.as_tensor(self, device="cuda")
        t = t.view(dtype).view(shape)
        t._xaac_holder = self  # keeps the mapping alive as long as the tensor
        return t

    def __del__(self):
        try:
            if self.ptr and self.ctx._h:
                if self.owned:
                    self.ctx._lib.xaac_b200_dev_free(self.ctx._h, ctypes.c_void_p(self.ptr))
                else:
                    self.ctx._lib.xaac_b200_ipc_close(self.ctx._h, ctypes.c_void_p(self.ptr))
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
        self.ptr = None
