package gammaspike.distribution;

import beast.base.core.Description;
import gammaspike.b200.GsmB200;
import gammaspike.b200.GsmContext;

/** BranchSpikePrior evaluated by libgsm_b200 (UNCOMPILED here); Inputs inherited (BranchSpikePrior.java:28-35). */
@Description("A sum of gamma distributions, one for each spike on a branch; evaluated by libgsm_b200")
public class BranchSpikePriorB200 extends BranchSpikePrior {
    private GsmContext ctx;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        if (treeInput.get() == null) throw new IllegalArgumentException("BranchSpikePriorB200 needs the tree input");
        ctx = GsmContext.of(treeInput.get());
        ctx.spikePrior = this;
    }

    /** BSP:53-187, both modes (known stub count BSP:77-100, mixture over the stub count BSP:102-172) */
    @Override
    public double calculateLogP() {
        logP = ctx.eval()[GsmB200.O_SPIKE_LOGP];
        return logP;
    }

    /** BSP:315-389; mu is recomputed on the device from the branch's heights exactly as Stubs.sampleNStubsOnBranch obtains it
     *  (Stubs:424-427 -> StumpedTreePrior.getMeanStubNumber STP:876-892), so the argument is not needed */
    @Override
    public double[] getCumulativeProbs(double mu, int nodeNr) { return ctx.stubCdf(nodeNr); }

    @Override
    protected boolean requiresRecalculation() {
        ctx.invalidate();
        return true;                                   // BSP:294-295
    }

    @Override
    public void restore() {
        ctx.invalidate();
        super.restore();
    }
}
