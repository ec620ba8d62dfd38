package gammaspike.b200;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.Map;
import java.util.WeakHashMap;

import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.Tree;
import beast.base.inference.parameter.RealParameter;
import gammaspike.clockmodel.PunctuatedRelaxedClockModel;
import gammaspike.distribution.BranchRatePrior;
import gammaspike.distribution.BranchSpikePrior;
import gammaspike.distribution.StumpedTreePrior;
import gammaspike.tree.Stubs;

import static gammaspike.b200.GsmB200.*;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * One gsm_handle per chain (keyed by the chain's Tree object) shared by the clock model, the three priors and the loggers:
 * whichever of them is asked first in an MCMC step uploads the state and launches ONE evaluation; the others read the same
 * result row.  The state is re-sent whole (heights, parents, rates, spikes, stub counts, 9 scalars: 36 bytes per node) -- a
 * single tree is far below any transfer threshold; batched drivers (many chains / proposals per call) use gsm_eval_host or
 * gsm_eval_dirty directly.  UNCOMPILED in the authoring environment.
 *
 * Staleness: BEAST calls requiresRecalculation() / store() / restore() on every CalculationNode; each B200 class forwards
 * them to {@link #invalidate()}, so a cached row never outlives the state it was computed from.
 */
public final class GsmContext implements AutoCloseable {
    private static final Map<Tree, GsmContext> CONTEXTS = new WeakHashMap<>();

    public static synchronized GsmContext of(Tree tree) {
        return CONTEXTS.computeIfAbsent(tree, GsmContext::new);
    }

    private final Tree tree;
    private final Arena arena = Arena.ofShared();
    private final MemorySegment handle;
    private final int n;
    private final MemorySegment height, parent, rate, spike, nstubs, params, useSpike, out, rates, wspikes;
    // the plugin objects of this chain register themselves (any of them may be absent)
    public PunctuatedRelaxedClockModel clock;
    public BranchSpikePrior spikePrior;
    public BranchRatePrior ratePrior;
    public StumpedTreePrior treePrior;
    private boolean valid = false, modeSet = false;
    private double[] row, ratesArr, wspikesArr;

    private GsmContext(Tree tree) {
        this.tree = tree;
        this.n = tree.getNodeCount();
        MemorySegment ph = arena.allocate(ADDRESS);
        check((int) call(CREATE, ph, Integer.getInteger("gsm.device", 0), 1L, n), MemorySegment.NULL);
        handle = ph.get(ADDRESS, 0);
        height = arena.allocate(JAVA_DOUBLE, n);
        parent = arena.allocate(JAVA_INT, n);
        rate = arena.allocate(JAVA_DOUBLE, n);
        spike = arena.allocate(JAVA_DOUBLE, n);
        nstubs = arena.allocate(JAVA_INT, n);
        params = arena.allocate(JAVA_DOUBLE, NPARAM);
        useSpike = arena.allocate(JAVA_BYTE, 1);
        out = arena.allocate(JAVA_DOUBLE, NOUT);
        rates = arena.allocate(JAVA_DOUBLE, n);
        wspikes = arena.allocate(JAVA_DOUBLE, n);
    }

    public MemorySegment handle() { return handle; }

    /** any StateNode of the chain changed, or the state was restored */
    public void invalidate() { valid = false; }

    /** Stubs.StubMode -> GSM_STUBS_* (Stubs.java:36-40, 643-650) */
    private static int stubMode(Stubs stubs) {
        if (stubs == null || !stubs.estimateStubs()) return STUBS_NOSTUB;
        return stubs.getReversibleJump() ? STUBS_EXACT : STUBS_COUNTS;
    }

    /** which optional Inputs are wired: the `input.get() != null` tests of PRCM:181-216, BSP:77, STP:174-219 */
    private void setMode() {
        int flags = 0;
        Stubs stubs = null;
        if (clock != null) {
            if (clock.indicatorInput.get() != null) flags |= F_HAS_INDICATOR;
            if (clock.noSpikeOnDatedTipsInput.get()) flags |= F_NO_SPIKE_DATED_TIPS;
            if (clock.ratesInput.get() != null) flags |= F_HAS_RATES;
            if (clock.relaxedInput.get() != null && !clock.relaxedInput.get().getValue()) flags |= F_RELAXED_FALSE;
            if (clock.stubsInput.get() != null) { flags |= F_STUBS_TO_CLOCK; stubs = clock.stubsInput.get(); }
        }
        if (spikePrior != null) {
            if (spikePrior.stubsInput.get() != null) { flags |= F_STUBS_TO_SPIKE_PRIOR; stubs = spikePrior.stubsInput.get(); }
            if (spikePrior.indicatorInput.get() != null) flags |= F_BSP_HAS_INDICATOR;
        }
        if (ratePrior != null) flags |= F_HAS_RATES;
        if (treePrior != null) {
            if (treePrior.stubsInput.get() != null) { flags |= F_STUBS_TO_TREE_PRIOR; stubs = treePrior.stubsInput.get(); }
            if (treePrior.ignoreTreePriorInput.get()) flags |= F_IGNORE_TREE_PRIOR;
            if (treePrior.ignoreStubPriorInput.get()) flags |= F_IGNORE_STUB_PRIOR;
            if (treePrior.conditionOnSamplingInput.get()) flags |= F_COND_SAMPLING;
            if (treePrior.conditionOnRhoSamplingInput.get()) flags |= F_COND_RHO_SAMPLING;
            if (treePrior.conditionOnRootInput.get()) flags |= F_COND_ROOT;
            if (treePrior.originInput.get() != null) flags |= F_ORIGIN;
        } else {
            flags |= F_IGNORE_TREE_PRIOR | F_IGNORE_STUB_PRIOR;
        }
        this.stubs = stubs;
        check((int) call(SET_MODE, handle, stubMode(stubs), flags), handle);
        modeSet = true;
    }
    private Stubs stubs;

    /** upload the current state and evaluate it, unless that has been done for this state already */
    public synchronized double[] eval() {
        if (valid) return row;
        if (!modeSet) setMode();
        for (Node nd : tree.getNodesAsArray()) {                               // Tree.getNodesAsArray(): index == getNr()
            height.setAtIndex(JAVA_DOUBLE, nd.getNr(), nd.getHeight());
            parent.setAtIndex(JAVA_INT, nd.getNr(), nd.isRoot() ? -1 : nd.getParent().getNr());
        }
        check((int) call(UPLOAD_TREES, handle, 1L, n, height, parent, MemorySegment.NULL), handle);
        RealParameter r = clock != null ? clock.ratesInput.get() : (ratePrior != null ? ratePrior.branchRatesInput.get() : null);
        RealParameter s = clock != null ? clock.spikesInput.get() : spikePrior.spikesInput.get();
        for (int i = 0; i < n; i++) {
            if (r != null) rate.setAtIndex(JAVA_DOUBLE, i, r.getValue(i));
            spike.setAtIndex(JAVA_DOUBLE, i, s.getValue(i));
        }
        final int mode = stubMode(stubs);
        if (mode == STUBS_COUNTS) {
            for (int i = 0; i < n; i++) nstubs.setAtIndex(JAVA_INT, i, stubs.getNStubsOnBranch(i));   // Stubs:352-355 (i >= dim -> 0)
        }
        check((int) call(UPLOAD_BRANCH, handle, r != null ? rate : MemorySegment.NULL, spike,
                         mode == STUBS_COUNTS ? nstubs : MemorySegment.NULL), handle);
        if (mode == STUBS_EXACT) {
            // entries 1 .. dim-1 of Stubs.branchNr / Stubs.time (index 0 is the dummy, Stubs:285-287)
            final int dim = stubs.getStubDimension();
            MemorySegment ptr = arena.allocate(JAVA_LONG, 2);
            ptr.setAtIndex(JAVA_LONG, 0, 0L);
            ptr.setAtIndex(JAVA_LONG, 1, Math.max(0, dim - 1));
            MemorySegment br = arena.allocate(JAVA_INT, Math.max(1, dim)), tm = arena.allocate(JAVA_DOUBLE, Math.max(1, dim));
            for (int k = 1; k < dim; k++) {
                br.setAtIndex(JAVA_INT, k - 1, stubs.getBranches().getValue(k));
                tm.setAtIndex(JAVA_DOUBLE, k - 1, stubs.getRelativeTimeOfStub(k));
            }
            check((int) call(UPLOAD_STUBS, handle, ptr, br, tm), handle);
        }
        double[] p = new double[NPARAM];
        p[P_CLOCK_RATE] = clock != null ? clock.meanRateInput.get().getArrayValue() : 1.0;          // PRCM:228
        p[P_SPIKE_MEAN] = clock != null ? clock.spikeMeanInput.get().getValue() : 1.0;              // PRCM:204
        p[P_SPIKE_SHAPE] = spikePrior != null ? spikePrior.shapeInput.get().getValue() : 1.0;       // BSP:58
        p[P_CLOCK_SD] = ratePrior != null ? ratePrior.sigmaInput.get().getValue() : 1.0;            // BRP:33
        if (treePrior != null) {                                                                     // STP:672-737 stay in Java
            p[P_LAMBDA] = treePrior.getLambda();
            p[P_MU] = treePrior.getMu();
            p[P_PSI] = treePrior.getPsi();
            p[P_RHO] = treePrior.getRho();
            p[P_ORIGIN] = treePrior.originInput.get() != null ? treePrior.originInput.get().getValue() : 0.0;
        } else {
            p[P_LAMBDA] = 1.0; p[P_MU] = 0.0; p[P_PSI] = 0.0; p[P_RHO] = 1.0;
        }
        MemorySegment.copy(p, 0, params, JAVA_DOUBLE, 0, NPARAM);
        boolean ind = clock == null || clock.indicatorInput.get() == null || clock.indicatorInput.get().getValue();
        useSpike.set(JAVA_BYTE, 0, (byte) (ind ? 1 : 0));
        check((int) call(UPLOAD_TREE_PARAMS, handle, params, useSpike), handle);
        check((int) call(EVAL, handle, out, rates, wspikes), handle);
        row = toArray(out, NOUT);
        ratesArr = toArray(rates, n);
        wspikesArr = toArray(wspikes, n);
        valid = true;
        return row;
    }

    public double[] rates() { eval(); return ratesArr; }       // PRCM.getRatesArray
    public double[] wspikes() { eval(); return wspikesArr; }   // SpikeSize.getArrayValue / PRCM.getBurstSize(int)

    /** BranchSpikePrior.getCumulativeProbs for one branch of the current state (BSP:315-389) */
    public double[] stubCdf(int nodeNr) {
        eval();
        try (Arena a = Arena.ofConfined()) {
            int cap = 4096;
            MemorySegment cdf = a.allocate(JAVA_DOUBLE, cap), len = a.allocate(JAVA_INT);
            check((int) call(STUB_CDF, handle, 0L, nodeNr, cdf, cap, len), handle);
            return toArray(cdf, Math.min(cap, len.get(JAVA_INT, 0)));
        }
    }

    /** Stubs.sampleNStubsOnBranch for every branch given one uniform per branch (Stubs:410-458) */
    public int[] stubChoose(double[] u) {
        eval();
        try (Arena a = Arena.ofConfined()) {
            MemorySegment us = a.allocate(JAVA_DOUBLE, n), res = a.allocate(JAVA_INT, n);
            MemorySegment.copy(u, 0, us, JAVA_DOUBLE, 0, n);
            check((int) call(STUB_CHOOSE, handle, 1L, MemorySegment.NULL, us, res), handle);
            return res.toArray(JAVA_INT);
        }
    }

    @Override
    public void close() {
        call(DESTROY, handle);
        arena.close();
    }
}
