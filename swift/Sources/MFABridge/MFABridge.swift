// MFABridge.swift -- Swift-facing forwarding layer over the CUDA build of MFAFFI.
//
// The reference's Sources/MFABridge/MFABridge.swift (782-805: context singleton, 1476-1550: forward entry points) implements
// every mfa_* symbol in Swift on top of Metal.  On B200 those symbols live in libMFAFFI.so, so nothing is re-implemented
// here: Swift callers import this module, keep handle-based calls, and every function forwards to the C ABI.
import CMFACuda

public enum MFAError: Error {
    case code(mfa_error_t)
}

@inline(__always)
private func check(_ rc: mfa_error_t) throws {
    if rc != MFA_SUCCESS { throw MFAError.code(rc) }
}

/// One retained context per device and process, as in the reference (MFABridge.swift:652-687).
public final class MFAContext {
    public let handle: mfa_context_t

    public init(device: Int32 = 0) throws {
        try check(mfa_set_device(device))
        var h: mfa_context_t? = nil
        try check(mfa_create_context(&h))
        handle = h!
    }

    deinit { mfa_destroy_context(handle) }

    public var lastGPULatency: Double { mfa_get_gpu_latency(handle) }
}

/// Zero-copy view of caller memory (host arrays are mirrored on the device for the duration of a call).
public final class MFABuffer {
    public let handle: mfa_buffer_t

    public init(context: MFAContext, pointer: UnsafeMutableRawPointer, byteCount: Int) throws {
        var h: mfa_buffer_t? = nil
        try check(mfa_buffer_from_ptr(context.handle, pointer, byteCount, &h))
        handle = h!
    }

    deinit { mfa_destroy_buffer(handle) }
}

/// O = softmax(scale * Q K^T [causal]) V; the precisions are the header's enum values (fp32 default, like the reference's strings).
public func attentionForward(_ ctx: MFAContext, q: MFABuffer, k: MFABuffer, v: MFABuffer, out: MFABuffer,
                             batch: UInt32, seqQ: UInt32, seqKV: UInt32, heads: UInt32, headDim: UInt16,
                             scale: Float, causal: Bool,
                             inputPrecision: mfa_precision_t = MFA_PRECISION_FP32,
                             outputPrecision: mfa_precision_t = MFA_PRECISION_FP32) throws {
    try check(mfa_attention_forward(ctx.handle, q.handle, k.handle, v.handle, out.handle,
                                    batch, seqQ, seqKV, heads, headDim, scale, causal,
                                    inputPrecision, MFA_PRECISION_FP32, outputPrecision,
                                    false, false, false, false,
                                    nil, 0, nil, nil, 0, MFA_MASK_TYPE_NONE, MFA_MASK_SCALAR_BYTE))
}

/// dQ, dK, dV from the saved O and L (log2-domain logsumexp), fp32 gradients.
public func attentionBackward(_ ctx: MFAContext, q: MFABuffer, k: MFABuffer, v: MFABuffer, out: MFABuffer,
                              gradOut: MFABuffer, lse: MFABuffer, gradQ: MFABuffer, gradK: MFABuffer, gradV: MFABuffer,
                              batch: UInt32, seqQ: UInt32, seqKV: UInt32, heads: UInt32, headDim: UInt16,
                              scale: Float, causal: Bool,
                              inputPrecision: mfa_precision_t = MFA_PRECISION_FP32) throws {
    try check(mfa_attention_backward(ctx.handle, gradOut.handle, q.handle, k.handle, v.handle, out.handle, lse.handle,
                                     gradQ.handle, gradK.handle, gradV.handle, nil,
                                     batch, seqQ, seqKV, heads, headDim, scale, causal, inputPrecision, MFA_PRECISION_FP32,
                                     false, false, false, false))
}
