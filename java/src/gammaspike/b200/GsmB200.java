package gammaspike.b200;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama (java.lang.foreign, JDK 22+) binding of include/gsm_b200.h, one static handle per C entry point.
 * UNCOMPILED in the authoring environment (no JDK).  No CPU fallback: a missing library fails class initialisation.
 */
public final class GsmB200 {
    private GsmB200() { }

    // ---- constants of gsm_b200.h ----
    public static final int ABI_VERSION = 2;
    public static final int F_HAS_INDICATOR = 1, F_NO_SPIKE_DATED_TIPS = 1 << 1, F_HAS_RATES = 1 << 2, F_RELAXED_FALSE = 1 << 3,
            F_STUBS_TO_CLOCK = 1 << 4, F_STUBS_TO_SPIKE_PRIOR = 1 << 5, F_STUBS_TO_TREE_PRIOR = 1 << 6, F_IGNORE_TREE_PRIOR = 1 << 7,
            F_IGNORE_STUB_PRIOR = 1 << 8, F_COND_SAMPLING = 1 << 9, F_COND_RHO_SAMPLING = 1 << 10, F_COND_ROOT = 1 << 11,
            F_ORIGIN = 1 << 12, F_BSP_HAS_INDICATOR = 1 << 13;
    public static final int STUBS_NOSTUB = 0, STUBS_COUNTS = 1, STUBS_EXACT = 2;
    public static final int NPARAM = 9, P_CLOCK_RATE = 0, P_SPIKE_MEAN = 1, P_SPIKE_SHAPE = 2, P_CLOCK_SD = 3, P_LAMBDA = 4, P_MU = 5,
            P_PSI = 6, P_RHO = 7, P_ORIGIN = 8;
    public static final int NOUT = 8, O_STP_LOGP = 0, O_TREE_LOGP = 1, O_STUB_LOGP = 2, O_SPIKE_LOGP = 3, O_RATE_LOGP = 4,
            O_SUM_BURST = 5, O_SUM_DIST = 6, O_N_EXTANT = 7;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.mapLibraryName("gsm_b200"), Arena.global());

    private static MethodHandle mh(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }
    private static FunctionDescriptor fi(java.lang.foreign.MemoryLayout... args) { return FunctionDescriptor.of(JAVA_INT, args); }

    public static final MethodHandle ABI = mh("gsm_abi_version", FunctionDescriptor.of(JAVA_INT));
    public static final MethodHandle DEVICE_COUNT = mh("gsm_device_count", FunctionDescriptor.of(JAVA_INT));
    public static final MethodHandle CREATE = mh("gsm_create", fi(ADDRESS, JAVA_INT, JAVA_LONG, JAVA_INT));
    public static final MethodHandle DESTROY = mh("gsm_destroy", fi(ADDRESS));
    public static final MethodHandle LAST_ERROR = mh("gsm_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    public static final MethodHandle SET_MODE = mh("gsm_set_mode", fi(ADDRESS, JAVA_INT, JAVA_INT));
    public static final MethodHandle SET_OPTION = mh("gsm_set_option", fi(ADDRESS, ADDRESS, JAVA_LONG));
    public static final MethodHandle UPLOAD_TREES = mh("gsm_upload_trees", fi(ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle UPLOAD_BRANCH = mh("gsm_upload_branch_params", fi(ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle UPLOAD_TREE_PARAMS = mh("gsm_upload_tree_params", fi(ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle UPLOAD_STUBS = mh("gsm_upload_stubs", fi(ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle EVAL = mh("gsm_eval", fi(ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle EVAL_HOST = mh("gsm_eval_host", fi(ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    /** h, n_props, tree_of_prop, dirty_ptr, dirty_nodes, new_height, new_rate, new_spike, new_nstubs, height_scale, new_params,
     *  new_use_spike, out_tree */
    public static final MethodHandle EVAL_DIRTY = mh("gsm_eval_dirty", fi(ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle STUB_CDF = mh("gsm_stub_cdf", fi(ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    public static final MethodHandle STUB_CDF_TREES = mh("gsm_stub_cdf_trees", fi(ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    public static final MethodHandle STUB_CHOOSE = mh("gsm_stub_choose", fi(ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle FORMAT_NEWICK = mh("gsm_format_newick", FunctionDescriptor.of(JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG));
    public static final MethodHandle PARSE_NEWICK = mh("gsm_parse_newick", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    public static final MethodHandle EFFECTIVE_LEAF_RATES = mh("gsm_effective_leaf_rates", fi(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle JAVA_DOUBLE_TO_STRING = mh("gsm_java_double_to_string", FunctionDescriptor.of(JAVA_LONG, JAVA_DOUBLE, ADDRESS, JAVA_LONG));
    public static final MethodHandle HOST_ALLOC = mh("gsm_host_alloc", fi(ADDRESS, JAVA_LONG));
    public static final MethodHandle HOST_FREE = mh("gsm_host_free", fi(ADDRESS));
    public static final MethodHandle DEVICE_STRIDE = mh("gsm_device_stride", FunctionDescriptor.of(JAVA_LONG, JAVA_INT));
    public static final MethodHandle BIND_DEVICE = mh("gsm_bind_device", fi(ADDRESS, JAVA_LONG, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle BIND_DEVICE_STUBS = mh("gsm_bind_device_stubs", fi(ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle EVAL_DEVICE = mh("gsm_eval_device", fi(ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle SYNC = mh("gsm_sync", fi(ADDRESS));
    public static final MethodHandle DEVICE_ALLOC = mh("gsm_device_alloc", fi(ADDRESS, JAVA_INT, JAVA_LONG));
    public static final MethodHandle DEVICE_FREE = mh("gsm_device_free", fi(ADDRESS, JAVA_INT));
    public static final MethodHandle IPC_EXPORT = mh("gsm_ipc_export", fi(ADDRESS, ADDRESS));
    public static final MethodHandle IPC_OPEN = mh("gsm_ipc_open", fi(JAVA_INT, ADDRESS, ADDRESS));
    public static final MethodHandle IPC_CLOSE = mh("gsm_ipc_close", fi(ADDRESS));
    public static final MethodHandle COPY_ASYNC = mh("gsm_copy_async", fi(ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    public static final MethodHandle MEASURE_FP64 = mh("gsm_measure_fp64_peak", fi(ADDRESS, ADDRESS));
    public static final MethodHandle LAUNCH_COUNT = mh("gsm_launch_count", FunctionDescriptor.of(JAVA_LONG, ADDRESS));
    public static final MethodHandle LAST_DIRTY_FULL = mh("gsm_last_dirty_full_count", FunctionDescriptor.of(JAVA_LONG, ADDRESS));

    static {
        if ((int) call(ABI) != ABI_VERSION) throw new UnsatisfiedLinkError("libgsm_b200 ABI version mismatch");
        if ((int) call(DEVICE_COUNT) < 1) throw new IllegalStateException("gsm_b200: no CUDA device; there is no CPU fallback");
    }

    public static Object call(MethodHandle h, Object... a) {
        try {
            return h.invokeWithArguments(a);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** throws with gsm_last_error's text when rc != GSM_OK */
    public static void check(int rc, MemorySegment handle) {
        if (rc == 0) return;
        MemorySegment msg = (MemorySegment) call(LAST_ERROR, handle);
        throw new IllegalStateException("gsm_b200 error " + rc + ": " + msg.reinterpret(1024).getString(0));
    }

    public static double[] toArray(MemorySegment s, int n) { return s.asSlice(0, 8L * n).toArray(JAVA_DOUBLE); }
}
