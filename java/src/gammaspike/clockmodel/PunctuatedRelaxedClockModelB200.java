package gammaspike.clockmodel;

import beast.base.core.Description;
import beast.base.evolution.tree.Node;
import gammaspike.b200.GsmContext;

/** PunctuatedRelaxedClockModel with its per-branch arithmetic on the B200 (UNCOMPILED here).  Inputs, ids and
 *  initAndValidate are inherited unchanged (PunctuatedRelaxedClockModel.java:29-159). */
@Description("Gamma spike relaxed clock; branch-rate map evaluated by libgsm_b200")
public class PunctuatedRelaxedClockModelB200 extends PunctuatedRelaxedClockModel {
    private GsmContext ctx;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        ctx = GsmContext.of(treeInput.get());
        ctx.clock = this;
    }

    /** PRCM:225-242: one array read; the whole tree was evaluated once for this MCMC state */
    @Override
    public double getRateForBranch(Node node) { return ctx.rates()[node.getNr()]; }

    /** PRCM:267-276 */
    @Override
    public double[] getRatesArray() { return ctx.rates().clone(); }

    /** PRCM:167-208 (GSMweightedSpikes, SpikeSize.java:48-56) */
    @Override
    public double getBurstSize(int nodeNr) { return ctx.wspikes()[nodeNr]; }

    @Override
    public double getBurstSize(Node node) { return ctx.wspikes()[node.getNr()]; }

    @Override
    protected boolean requiresRecalculation() {
        ctx.invalidate();
        return super.requiresRecalculation();
    }

    @Override
    protected void restore() {
        ctx.invalidate();
        super.restore();
    }
}
