package gammaspike.distribution;

import beast.base.core.Description;
import beast.base.evolution.tree.Tree;
import beast.base.evolution.tree.TreeInterface;
import gammaspike.b200.GsmB200;
import gammaspike.b200.GsmContext;

/** StumpedTreePrior (BDWS / FBDWS, rho-sampling, origin / root conditioning, stub densities of all three Stubs modes)
 *  evaluated by libgsm_b200 (UNCOMPILED here).  Inputs and their validation are inherited (StumpedTreePrior.java:24-50,
 *  85-108); the parameter maps getLambda / getMu / getPsi / getRho (STP:672-737) are the inherited Java methods. */
@Description("Stumped tree prior; evaluated by libgsm_b200")
public class StumpedTreePriorB200 extends StumpedTreePrior {
    private GsmContext ctx;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        ctx = GsmContext.of((Tree) treeInput.get());
        ctx.treePrior = this;
    }

    /** STP:150-284.  The device returns treeLogP + stubLogP and the number of extant leaves; the cross-call nExtantTaxa
     *  guard (STP:196-200) is host state and stays here. */
    @Override
    public double calculateTreeLogLikelihood(final TreeInterface tree) {
        final double[] row = ctx.eval();
        final boolean conditioned = originInput.get() != null || conditionOnRootInput.get();       // STP:179-186 skip the guard
        if (!conditioned && getLambda() > getMu() && !ignoreTreePriorInput.get()) {
            final int n = (int) row[GsmB200.O_N_EXTANT];
            if (initialising) nExtantTaxa = n;                                                     // STP:196-200
            else if (n != nExtantTaxa) return Double.NEGATIVE_INFINITY;
        }
        initialising = false;
        logP = row[GsmB200.O_STP_LOGP];
        return logP;
    }

    @Override
    protected boolean requiresRecalculation() {
        ctx.invalidate();
        return true;                                   // STP:836-842
    }

    @Override
    public void restore() {
        ctx.invalidate();
        super.restore();
    }
}
