"""Drop-in for the reference's ``ldpc_jossy/py/ldpc.py``: the ``code`` object.

Same constructor, attributes, method names, argument meaning, return types and
``NameError`` conditions as the reference (ldpc.py:4-503), with the decoder
running on the GPU through libldpcsparc_b200.so:

* ``decode(ch)`` with ``ch`` of shape ``[N]`` behaves like the reference
  (returns ``(app float64 ndarray[N], it int)``);
* ``decode(ch)`` with ``ch`` of shape ``[F, N]`` decodes F independent frames in
  one call and returns ``(app[F, N], it ndarray[F])`` -- the batched axis the
  reference's Monte-Carlo loops (ldpc_awgn.py:88-107) iterate over serially;
* a CUDA ``torch.Tensor`` input stays on the device and returns tensors.

Arithmetic is fp32 by default (``precision='fp32'``, the throughput mode;
LLRs within 1e-4 relative of the reference) or fp64 (``precision='fp64'``, the
parity mode).  There is no CPU decoder here: without the CUDA library / a GPU
``decode`` raises.
"""
import ctypes as ct

import numpy as np

from . import _lib
from .protographs import TABLES

DEFAULT_PRECISION = "fp32"


def set_default_precision(p):
    global DEFAULT_PRECISION
    if p not in ("fp32", "fp64"):
        raise ValueError("precision must be 'fp32' or 'fp64'")
    DEFAULT_PRECISION = p


def _parse_table(rows):
    return np.array([[-1 if t == "." else int(t) for t in row.split()] for row in rows])


class code:
    def __init__(self, standard="802.11n", rate="1/2", z=27, ptype="A"):
        self.standard = standard
        self.rate = rate
        self.z = z
        self.ptype = ptype
        self.proto = self.assign_proto()
        vdeg, cdeg, intrlv = self.prepare_decoder()
        self.vdeg = vdeg
        self.cdeg = cdeg
        self.intrlv = intrlv
        self.Nv = len(vdeg)
        self.Nc = len(cdeg)
        self.Nmsg = len(intrlv)
        self.N = self.Nv
        self.K = self.Nv - self.Nc
        self._graph = None   # device handle, created on first decode

    # ------------------------------------------------------------ protograph
    def assign_proto(self):
        """Protograph of the selected IEEE code (reference ldpc.py:24-272).

        Raises NameError for the same conditions and with the same messages."""
        standard, rate, z, ptype = self.standard, self.rate, self.z, self.ptype
        if standard == "802.16":
            if rate in ("1/2", "5/6"):
                key = (standard, rate, None, None)
            elif rate in ("2/3", "3/4"):
                if ptype not in ("A", "B"):
                    raise NameError("802.16 type must be either A or B")
                key = (standard, rate, None, ptype)
            else:
                raise NameError("802.16 invalid rate")
        elif standard == "802.11n":
            if z not in (27, 54, 81):
                raise NameError("802.11n invalid z (must be 27,54 or 81)")
            if rate not in ("1/2", "2/3", "3/4", "5/6"):
                raise NameError("802.11n invalid rate")
            key = (standard, rate, z, None)
        else:
            raise NameError("IEEE standard unknown")
        return _parse_table(TABLES[key])

    def pcmat(self):
        """Parity-check matrix (reference ldpc.py:275-300)."""
        proto, z = self.proto, self.z
        H = np.zeros((z * proto.shape[0], z * proto.shape[1]), dtype=int)
        k = np.arange(z)
        for r, c in zip(*np.nonzero(proto != -1)):
            H[r * z + k, c * z + (k + proto[r, c]) % z] = 1
        return H

    def prepare_decoder(self):
        """vdeg, cdeg, intrlv of the expanded graph (reference ldpc.py:303-396).

        Same numbering as the reference: messages are stored check by check with
        a check's ports in increasing protograph column; a variable's ports are
        in increasing check (protograph row) order; ``intrlv[variable port]`` is
        the message slot.  Built in closed form instead of the reference's port
        search (it visits protograph entries row-major, so the n-th entry met in
        a row / column gets that row's / column's n-th port)."""
        proto, z = self.proto, self.z
        mask = proto != -1
        cdeg = np.repeat(np.sum(mask, 1), z)
        vdeg = np.repeat(np.sum(mask, 0), z)
        cum_c = np.concatenate(([0], np.cumsum(cdeg)))
        cum_v = np.concatenate(([0], np.cumsum(vdeg)))
        rows, cols = np.nonzero(mask)                       # row-major
        cport = (np.cumsum(mask, 1) - 1)[rows, cols]        # rank of the entry in its row
        vport = (np.cumsum(mask, 0) - 1)[rows, cols]        # rank of the entry in its column
        k = np.arange(z)
        cind = rows[:, None] * z + k[None, :]
        vind = cols[:, None] * z + (k[None, :] + proto[rows, cols][:, None]) % z
        slot = cum_c[cind] + cport[:, None]
        vslot = cum_v[vind] + vport[:, None]
        intrlv = np.empty(int(cum_c[-1]), dtype=np.int32)   # int32 as the reference (ldpc.py:395)
        intrlv[vslot.ravel()] = slot.ravel()
        return vdeg, cdeg, intrlv

    # --------------------------------------------------------------- encoder
    def encode(self, info):
        """Systematic encoder (reference ldpc.py:400-460).  ``info`` is ``[K]``
        (returns ``[N]``) or ``[F, K]`` (returns ``[F, N]``)."""
        z, proto = self.z, self.proto
        Np, Mp = proto.shape[1], proto.shape[0]
        Kp = Np - Mp
        K = Kp * z
        info = np.asarray(info)
        single = info.ndim == 1
        if info.shape[-1] != K or info.ndim > 2:
            raise NameError("information word length not compatible with proto and z")
        u = info.reshape(-1, Kp, z).astype(np.int64)
        F = u.shape[0]
        x = np.zeros((F, Np, z), dtype=np.int64)
        x[:, :Kp] = u
        p = np.zeros((F, Mp, z), dtype=np.int64)
        for j in range(Mp):
            for k in np.nonzero(proto[j, :Kp] != -1)[0]:
                p[:, j] += np.roll(x[:, k], -proto[j, k], axis=-1)
        p %= 2
        tp = p.sum(1) % 2
        toff = np.zeros(z, dtype=int)
        for j in np.nonzero(proto[:, Kp] != -1)[0]:
            toff[proto[j, Kp] % z] += 1
        tnz = np.nonzero(toff % 2)[0]
        if len(tnz) != 1:
            raise NameError("The offsets in colum Kp+1 of proto do not add to a single offset")
        x[:, Kp] = np.roll(tp, tnz[0], axis=-1)
        for j in range(Mp - 1):
            myk = Kp + j + 1
            acc = p[:, j].copy()
            for k in np.nonzero(proto[j, Kp:myk] != -1)[0]:
                acc += np.roll(x[:, Kp + k], -proto[j, Kp + k], axis=-1)
            x[:, myk] = acc % 2
        x = x.reshape(F, Np * z)
        return x[0] if single else x

    # --------------------------------------------------------------- decoder
    def _handle(self):
        if self._graph is None:
            _lib.require_device()
            L = _lib.lib()
            h = ct.c_void_p()
            proto = np.ascontiguousarray(self.proto, dtype=np.int32)
            rc = L.lsb_ldpc_graph_create_qc(
                proto.ctypes.data_as(_lib.c_i32p), proto.shape[0], proto.shape[1], int(self.z), ct.byref(h)
            )
            if rc == -1 or rc == -4:
                # protograph outside the QC kernel's table limits: generic graph
                vd = np.ascontiguousarray(self.vdeg, dtype=np.int64)
                cd = np.ascontiguousarray(self.cdeg, dtype=np.int64)
                il = np.ascontiguousarray(self.intrlv, dtype=np.int64)
                rc = L.lsb_ldpc_graph_create(
                    vd.ctypes.data_as(_lib.c_i64p), cd.ctypes.data_as(_lib.c_i64p), il.ctypes.data_as(_lib.c_i64p),
                    self.Nv, self.Nc, self.Nmsg, ct.byref(h),
                )
            _lib.check(rc, "lsb_ldpc_graph_create")
            self._graph = h
        return self._graph

    def __del__(self):
        try:
            if getattr(self, "_graph", None) is not None:
                _lib.lib().lsb_ldpc_graph_destroy(self._graph)
                self._graph = None
        except Exception:
            pass

    def decode(self, ch, max_itcount=200, dectype="sumprod2", corr_factor=0.7, precision=None):
        """Belief-propagation decoding (reference ldpc.py:463-490).

        ch: channel LLRs, ``[N]`` or ``[F, N]`` (numpy, or a CUDA torch tensor).
        Returns ``(app, it)``: a-posteriori LLRs (float64 for numpy input, as the
        reference) and the iteration count (int for one frame, int32 array for a
        batch).  ``dectype`` in {'sumprod', 'sumprod2', 'minsum'}."""
        if isinstance(max_itcount, str):
            # pre-fork call style decode(y, 'sumprod') still used by the
            # reference's own test (test_ldpc.py:62,66)
            max_itcount, dectype = 200, max_itcount
        try:
            import torch
            is_tensor = isinstance(ch, torch.Tensor)
        except ImportError:   # pragma: no cover
            is_tensor = False
        n_last = ch.shape[-1] if hasattr(ch, "shape") and len(ch.shape) else len(ch)
        if n_last != len(self.vdeg):
            raise NameError("Channel inputs not consistent with variable degrees")
        if dectype not in _lib.ALGO:
            raise NameError("Decoder type unknonwn")
        algo = _lib.ALGO[dectype]
        precision = precision or DEFAULT_PRECISION
        if precision not in ("fp32", "fp64"):
            raise ValueError("precision must be 'fp32' or 'fp64'")
        if is_tensor:
            return self._decode_tensor(ch, int(max_itcount), algo, float(corr_factor), precision)

        ch = np.asarray(ch)
        single = ch.ndim == 1
        if ch.ndim > 2:
            raise NameError("Channel inputs not consistent with variable degrees")
        g = self._handle()
        L = _lib.lib()
        F = 1 if single else ch.shape[0]
        iters = np.zeros(F, dtype=np.int32)
        if ch.dtype == np.float32 and precision == "fp32":
            llr = np.ascontiguousarray(ch.reshape(F, self.Nv))
            app = np.zeros((F, self.Nv), dtype=np.float32)
            rc = L.lsb_ldpc_decode_host_f32(g, llr.ctypes.data, app.ctypes.data, None, iters.ctypes.data,
                                            F, int(max_itcount), algo, float(corr_factor))
            app = app.astype(np.float64)
        else:
            llr = np.ascontiguousarray(ch.reshape(F, self.Nv), dtype=np.float64)
            app = np.zeros((F, self.Nv), dtype=np.float64)
            rc = L.lsb_ldpc_decode_host_f64(g, llr.ctypes.data, app.ctypes.data, None, iters.ctypes.data,
                                            F, int(max_itcount), algo, float(corr_factor),
                                            1 if precision == "fp64" else 0)
        _lib.check(rc, "lsb_ldpc_decode_host")
        if single:
            return app[0], int(iters[0])
        return app, iters

    def decode_bits(self, ch, max_itcount=200, corr_factor=0.7):
        """Host-array Monte-Carlo entry (``lsb_ldpc_decode_host_bits_f32``): numpy float32 LLRs ``[F, N]`` in,
        ``(bits uint32 [F, ceil(N / 32)], it int32 [F])`` out, sum-product (`sumprod2`) in fp32.  Bit ``v % 32`` of
        word ``v // 32`` of a frame is ``app[v] < 0`` -- what the reference's loop keeps of a decode
        (ldpc_awgn.py:97) -- so 1/32 of the APP bytes come back to the host; `unpack_bits` undoes the packing."""
        ch = np.asarray(ch)
        if ch.ndim != 2 or ch.shape[1] != len(self.vdeg):
            raise NameError("Channel inputs not consistent with variable degrees")
        llr = np.ascontiguousarray(ch, dtype=np.float32)
        F = llr.shape[0]
        bits = np.zeros((F, (self.Nv + 31) // 32), dtype=np.uint32)
        iters = np.zeros(F, dtype=np.int32)
        rc = _lib.lib().lsb_ldpc_decode_host_bits_f32(self._handle(), llr.ctypes.data, bits.ctypes.data, iters.ctypes.data,
                                                      F, int(max_itcount), _lib.ALGO["sumprod2"], float(corr_factor))
        _lib.check(rc, "lsb_ldpc_decode_host_bits")
        return bits, iters

    def unpack_bits(self, bits):
        """``[F, ceil(N / 32)]`` words of `decode_bits` -> uint8 ``[F, N]`` hard decisions."""
        bits = np.asarray(bits, dtype=np.uint32)
        out = (bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
        return out.reshape(bits.shape[0], -1)[:, :self.Nv].astype(np.uint8)

    def _decode_tensor(self, ch, max_itcount, algo, corr_factor, precision):
        import torch
        if not ch.is_cuda:
            raise TypeError("torch input must be a CUDA tensor (numpy arrays take the host path)")
        single = ch.dim() == 1
        want = torch.float64 if precision == "fp64" else torch.float32
        llr = ch.reshape(-1, self.Nv).to(want).contiguous()
        F = llr.shape[0]
        with torch.cuda.device(llr.device):
            g = self._handle()
            app = torch.empty_like(llr)
            iters = torch.empty(F, dtype=torch.int32, device=llr.device)
            fn = _lib.lib().lsb_ldpc_decode_dev_f64 if precision == "fp64" else _lib.lib().lsb_ldpc_decode_dev_f32
            rc = fn(g, llr.data_ptr(), app.data_ptr(), None, iters.data_ptr(), F, max_itcount, algo,
                    corr_factor, torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "lsb_ldpc_decode_dev")
        if single:
            return app[0], int(iters[0].item())
        return app, iters

    def decode_hard(self, llr, max_itcount=200, dectype="sumprod2", corr_factor=0.7, precision=None, out=None):
        """Batched Monte-Carlo entry: CUDA tensor ``llr [F, N]`` in, ``(hard uint8
        [F, N], it int32 [F])`` out, nothing else written.  (New API; the
        reference computes ``app < 0`` on the host, ldpc_awgn.py:97.)"""
        import torch
        if dectype not in _lib.ALGO:
            raise NameError('Decoder type unknown')                 # same conditions as decode (ldpc.py:470,489)
        if llr.dim() != 2 or llr.shape[1] != self.Nv:
            raise NameError('Channel inputs not consistent with variable degrees')
        precision = precision or DEFAULT_PRECISION
        want = torch.float64 if precision == "fp64" else torch.float32
        if llr.dtype != want or not llr.is_contiguous():
            llr = llr.to(want).contiguous()
        F = llr.shape[0]
        with torch.cuda.device(llr.device):
            g = self._handle()
            if out is None:
                hard = torch.empty((F, self.Nv), dtype=torch.uint8, device=llr.device)
                iters = torch.empty(F, dtype=torch.int32, device=llr.device)
            else:
                hard, iters = out
            fn = _lib.lib().lsb_ldpc_decode_dev_f64 if precision == "fp64" else _lib.lib().lsb_ldpc_decode_dev_f32
            rc = fn(g, llr.data_ptr(), None, hard.data_ptr(), iters.data_ptr(), F, int(max_itcount),
                    _lib.ALGO[dectype], float(corr_factor), torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "lsb_ldpc_decode_dev")
        return hard, iters

    def mc_decode(self, codewords, n_frames, sigma2, seed=0, frame0=0, max_itcount=200, dectype="sumprod2",
                  corr_factor=0.7, counters=None, want_llr=False):
        """Fused Monte-Carlo step on the GPU (``lsb_ldpc_mc_dev_f32``): BPSK + AWGN(sigma2) + LLR scaling are
        generated inside the decoder kernel (Philox keyed by ``(seed, frame0 + f)``), the decode runs, and the
        wrong hard decisions over all N code bits are counted in its epilogue -- the body of the reference's loop
        (ldpc_awgn.py:88-107) without an LLR batch in HBM.  ``codewords``: CUDA uint8 ``[n_cw, N]`` pool; frame f
        carries codeword ``(frame0 + f) % n_cw``.  Returns ``(bit_errors int32 [F], iters int32 [F])`` (+ the
        generated LLRs ``[F, N]`` with ``want_llr``); ``counters`` (CUDA int64 [4], optional) accumulates
        bit errors, frame errors, frames, iterations."""
        import torch
        if not codewords.is_cuda or codewords.dtype != torch.uint8 or codewords.shape[-1] != self.N:
            raise TypeError("codewords must be a CUDA uint8 tensor [n_cw, N]")
        cw = codewords.reshape(-1, self.N).contiguous()
        F = int(n_frames)
        with torch.cuda.device(cw.device):
            g = self._handle()
            be = torch.empty(F, dtype=torch.int32, device=cw.device)
            it = torch.empty(F, dtype=torch.int32, device=cw.device)
            llr = torch.empty((F, self.N), dtype=torch.float32, device=cw.device) if want_llr else None
            rc = _lib.lib().lsb_ldpc_mc_dev_f32(g, cw.data_ptr(), cw.shape[0], float(sigma2), int(seed) & (2 ** 64 - 1),
                                                int(frame0), F, int(max_itcount), _lib.ALGO[dectype], float(corr_factor),
                                                be.data_ptr(), it.data_ptr(), llr.data_ptr() if want_llr else None,
                                                counters.data_ptr() if counters is not None else None,
                                                torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "lsb_ldpc_mc_dev_f32")
        return (be, it, llr) if want_llr else (be, it)

    def encode_device(self, info):
        """Systematic encoding on the GPU: CUDA tensor ``info [F, K]`` (any integer / bool
        dtype, values 0/1) -> ``uint8 [F, N]`` codewords (same map as ``encode``)."""
        import torch
        if not info.is_cuda:
            raise TypeError("encode_device takes a CUDA tensor (numpy arrays: use encode)")
        if info.shape[-1] != self.K:
            raise NameError('information word length not compatible with proto and z')
        u = info.reshape(-1, self.K).to(torch.uint8).contiguous()
        with torch.cuda.device(u.device):
            g = self._handle()
            x = torch.empty((u.shape[0], self.N), dtype=torch.uint8, device=u.device)
            _lib.check(_lib.lib().lsb_ldpc_encode_dev(g, u.data_ptr(), x.data_ptr(), u.shape[0],
                                                      torch.cuda.current_stream().cuda_stream), "lsb_ldpc_encode_dev")
        return x

    # -------------------------------------------------------- scalar helpers
    def Lxor(self, L1, L2, corrflag=1):
        """LLR of the XOR of two bits (reference ldpc.py:492-495, c_ldpc.c:234)."""
        _lib.require_device()
        return _lib.lib().Lxor(float(L1), float(L2), int(corrflag))

    def Lxfb(self, L, corrflag=1):
        """Extrinsic LLRs of a check node (reference ldpc.py:497-503, c_ldpc.c:294)."""
        _lib.require_device()
        L = np.array(L, dtype=float)
        out = _lib.lib().Lxfb(L.ctypes.data_as(_lib.c_f64p), len(L), int(corrflag))
        return out, L
